"""Command-line mirror of /root/reference/inference.py:41-150: two images -> Gaussians + panoptic labels -> output.ply.

  python -m siu3r_b200.inference --model_path ckpt --image_path1 a.jpg --image_path2 b.jpg --output_path out [--fx --fy --cx --cy]

Same arguments and defaults as the reference script.  The checkpoint is the reference's Lightning checkpoint (keys prefixed
"model.") or a plain state_dict; `--synthetic_weights` substitutes the seeded random weights used by the test-suite.
"""
from __future__ import annotations

from argparse import ArgumentParser
from pathlib import Path

import torch

from .io import default_intrinsics, export_ply, load_checkpoint, preprocess_image


def main(argv=None):
    ap = ArgumentParser()
    ap.add_argument("--model_path", type=str, default="pretrained_weights/siu3r_epoch100.ckpt")
    ap.add_argument("--image_path1", type=str, default="assets/living_room_image1.jpg")
    ap.add_argument("--image_path2", type=str, default="assets/living_room_image2.jpg")
    ap.add_argument("--output_path", type=str, default="infer_outputs")
    ap.add_argument("--cx", type=float, default=128.0)
    ap.add_argument("--cy", type=float, default=128.0)
    ap.add_argument("--fx", type=float, default=318.0)
    ap.add_argument("--fy", type=float, default=318.0)
    ap.add_argument("--precision", default="tf32", choices=["tf32", "fp32x3"])
    ap.add_argument("--synthetic_weights", action="store_true")
    ap.add_argument("--gpu_ingest", action="store_true", help="resize / crop the decoded frames on the GPU (bit-identical to the PIL recipe)")
    args = ap.parse_args(argv)
    out = Path(args.output_path)
    out.mkdir(parents=True, exist_ok=True)
    for p in (args.image_path1, args.image_path2):
        if not Path(p).exists():
            raise FileNotFoundError(f"Image file {p} does not exist.")
    if args.synthetic_weights:
        from .synth import make_state_dict
        sd = make_state_dict()
    else:
        if not Path(args.model_path).exists():
            raise FileNotFoundError(f"Model file {args.model_path} does not exist.")
        sd = load_checkpoint(args.model_path)
    from .model import ModelCfg, SIU3RModel
    if args.gpu_ingest:
        from .io import preprocess_views_cuda
        images = preprocess_views_cuda([args.image_path1, args.image_path2])                                                # [1, 2, 3, 256, 256] on the device
    else:
        images = torch.stack([preprocess_image(args.image_path1), preprocess_image(args.image_path2)], dim=0).unsqueeze(0)   # [1, 2, 3, 256, 256]
    intrinsics = default_intrinsics(args.fx, args.fy, args.cx, args.cy)
    model = SIU3RModel(ModelCfg(image_size=(256, 256)), precision=args.precision)
    model.load_state_dict(sd)
    model.cuda()
    g, seg_output, seg_masks, seg_infos, q_scores = model(images.cuda(), intrinsics.cuda(), enable_query_class_logit_lift=True)
    path = export_ply(means=g.means[0], scales=g.scales[0], rotations=g.rotations[0], harmonics=g.harmonics[0], opacities=g.opacities[0],
                      semantic_labels=g.semantic_labels[0], instance_labels=g.instance_labels[0], seg_query_class_logits=g.seg_query_class_logits[0],
                      path=out / "output.ply", shift_and_scale=False, save_sh_dc_only=False)
    print(f"wrote {path} ({path.stat().st_size / 1e6:.1f} MB), segments: {seg_infos[0]}")
    return path


if __name__ == "__main__":
    main()
