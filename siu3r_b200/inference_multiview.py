"""Command-line mirror of /root/reference/inference_multiview.py:40-153: a directory of V >= 2 images -> Gaussians + panoptic labels -> output.ply.

  python -m siu3r_b200.inference_multiview --model_path ckpt --image_dir assets/4views --output_path out [--fx --fy --cx --cy]

Same arguments and defaults as the reference script: the frames are the *.jpg, *.png and *.jpeg files of the directory (each group sorted, in
that order, :95-100), every view gets the same intrinsics (:108-116), the model is the V-view class (SIU3RMultiViewModel <->
PipelineMultiView.model) and the first file is the reference view.  `--synthetic_weights` substitutes the seeded random weights of the
test-suite, `--gpu_ingest` resizes / crops the decoded frames on the GPU (bit-identical to the PIL recipe).
"""
from __future__ import annotations

from argparse import ArgumentParser
from pathlib import Path

import torch

from .io import default_intrinsics, export_ply, load_checkpoint, preprocess_image


def main(argv=None):
    ap = ArgumentParser()
    ap.add_argument("--model_path", type=str, default="pretrained_weights/siu3r_4view.ckpt")
    ap.add_argument("--image_dir", type=str, default="assets/4views")
    ap.add_argument("--output_path", type=str, default="infer_outputs")
    ap.add_argument("--cx", type=float, default=128.0)
    ap.add_argument("--cy", type=float, default=128.0)
    ap.add_argument("--fx", type=float, default=318.0)
    ap.add_argument("--fy", type=float, default=318.0)
    ap.add_argument("--precision", default="tf32", choices=["tf32", "fp32x3"])
    ap.add_argument("--synthetic_weights", action="store_true")
    ap.add_argument("--gpu_ingest", action="store_true")
    args = ap.parse_args(argv)
    out = Path(args.output_path)
    out.mkdir(parents=True, exist_ok=True)
    if not args.synthetic_weights and not Path(args.model_path).exists():
        raise FileNotFoundError(f"Model file {args.model_path} does not exist.")
    images_dir = Path(args.image_dir)
    if not images_dir.exists():
        raise FileNotFoundError(f"Image directory {images_dir} does not exist.")
    image_paths = sorted(images_dir.glob("*.jpg")) + sorted(images_dir.glob("*.png")) + sorted(images_dir.glob("*.jpeg"))
    assert len(image_paths) >= 2, f"{images_dir} holds {len(image_paths)} image(s); the multi-view model needs at least two context views"
    if args.synthetic_weights:
        from .synth import make_state_dict
        sd = make_state_dict()
    else:
        sd = load_checkpoint(args.model_path)
    from .model import ModelCfg, SIU3RMultiViewModel
    if args.gpu_ingest:
        from .io import preprocess_views_cuda
        images = preprocess_views_cuda(image_paths)                                                   # [1, V, 3, 256, 256] on the device
    else:
        images = torch.stack([preprocess_image(p) for p in image_paths], dim=0).unsqueeze(0).cuda()  # [1, V, 3, 256, 256]
    V = images.shape[1]
    intrinsics = default_intrinsics(args.fx, args.fy, args.cx, args.cy, views=V)                     # identical K for every view (:108-116)
    model = SIU3RMultiViewModel(ModelCfg(image_size=(256, 256)), precision=args.precision)
    model.load_state_dict(sd)
    model.cuda()
    g, seg_output, seg_masks, seg_infos, q_scores = model(images, intrinsics.cuda(), enable_query_class_logit_lift=True)
    path = export_ply(means=g.means[0], scales=g.scales[0], rotations=g.rotations[0], harmonics=g.harmonics[0], opacities=g.opacities[0],
                      semantic_labels=g.semantic_labels[0], instance_labels=g.instance_labels[0], seg_query_class_logits=g.seg_query_class_logits[0],
                      path=out / "output.ply", shift_and_scale=False, save_sh_dc_only=False)
    print(f"wrote {path} ({path.stat().st_size / 1e6:.1f} MB), {V} views, segments: {seg_infos[0]}")
    return path


if __name__ == "__main__":
    main()
