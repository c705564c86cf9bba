"""ORACLE / TEST INFRASTRUCTURE: generate tests/golden/model_S*.npz from the UNMODIFIED reference (CPU, fp32).

Run in the build container (needs /root/reference):   python oracle/make_golden.py 64 256 512
                                                       python oracle/make_golden.py --views=4 64 256 512   (multi-view model)
                                                       python oracle/make_golden.py 64x96                   (rectangular H x W frame)
Each fixture holds, per stage boundary of SIU3RModel.forward, the tensor's shape / mean / abs-mean / abs-max and
2048 samples at fixed pseudo-random flat indices (oracle/ref_model.py:sample_indices), plus the full small outputs
(class logits, segment infos, label histograms).  Inputs and weights are regenerated on the GPU box from seeds
(siu3r_b200.synth), so only these summaries travel.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_model as R  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main(sizes, views=2):
    """views == 2: SIU3RModel (model_S*.npz); views > 2: SIU3RMultiViewModel (model_V{views}_S*.npz, BASELINE config 4)."""
    os.makedirs(OUT, exist_ok=True)
    sd = R.make_state_dict()
    for S in sizes:
        t0 = time.time()
        torch.set_num_threads(os.cpu_count())
        model = R.build_reference(S, sd, multiview=views > 2)
        img, K = R.synthetic_inputs(1, views, S)
        t1 = time.time()
        st = (R.run_reference_stages_multi if views > 2 else R.run_reference_stages)(model, img, K)
        t2 = time.time()
        arrays, meta = {}, {"size": S, "views": views, "forward_s": t2 - t1, "threads": torch.get_num_threads()}
        for name, v in st.items():
            if torch.is_tensor(v):
                sm = R.summarize(v)
                arrays[name + "__samples"] = sm.pop("samples")
                meta[name] = sm
        arrays["class_queries_logits__full"] = st["class_queries_logits"].numpy()
        meta["seg_infos"] = st["seg_infos"]
        meta["query_scores"] = st["query_scores"]
        sm = st["seg_masks"][0]
        arrays["seg_mask0__samples"] = R.summarize(sm)["samples"]
        meta["seg_mask0"] = {"shape": list(sm.shape), "dtype": str(sm.dtype), "hist": torch.bincount((sm.flatten().long() + 1)).tolist()}
        qc = st["seg_query_class_logits"][0]
        s2 = R.summarize(qc)
        arrays["qc0__samples"] = s2.pop("samples")
        meta["qc0"] = s2
        meta["sem_hist"] = torch.bincount(st["g_semantic_labels"].flatten().long(), minlength=22).tolist()
        meta["inst_hist"] = torch.bincount(st["g_instance_labels"].flatten().long()).tolist()
        tag = S if isinstance(S, int) else f"{S[0]}x{S[1]}"
        name = f"model_S{tag}.npz" if views == 2 else f"model_V{views}_S{tag}.npz"
        np.savez_compressed(os.path.join(OUT, name), meta=json.dumps(meta), **arrays)
        print(f"S={S}: build {t1 - t0:.1f}s forward {t2 - t1:.1f}s infos={st['seg_infos']}", flush=True)


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--views=")]
    nv = [int(a.split("=")[1]) for a in sys.argv[1:] if a.startswith("--views=")]
    main([tuple(int(x) for x in a.split("x")) if "x" in a else int(a) for a in args] or [64, 256], views=nv[0] if nv else 2)
