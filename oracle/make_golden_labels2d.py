"""Generates tests/golden/labels2d_cases.npz by running the reference's OWN statements for the 2-D label extraction.

The statements live inside Pipeline.validation-side code (/root/reference/src/pipeline.py, between the line that reads
`render_output["render_qc_logits"]` and the `torch.stack(all_sem_id, ...)` that follows the per-sample loop); the module cannot be
imported here (lightning is absent), so this script reads that span of the file AT GENERATION TIME, dedents it and executes it
unmodified with a stand-in `self` (device = cpu, pipecfg.model.mask2former.{label_ids_to_fuse, num_queries}).  Nothing of the
reference's text is stored in this repository; only the seeded inputs and the outputs are.

    python oracle/make_golden_labels2d.py            (needs /root/reference; CPU only)
"""
from __future__ import annotations

import json
import os
import sys
import textwrap
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src/pipeline.py"
START = 'render_qc_logits = render_output["render_qc_logits"]'
STOP = "all_sem_id = torch.stack(all_sem_id, dim=0)"

# name: (v, q, c, h, w, seed, kind)
CASES = {
    "small": (2, 5, 21, 12, 16, 0, "peaky"),
    "ties": (1, 4, 21, 8, 8, 1, "quantised"),      # many equal logits: exercises first-index tie breaking
    "onequery": (3, 1, 21, 6, 10, 2, "peaky"),
    "manyclasses": (1, 7, 40, 9, 7, 3, "peaky"),   # more classes than a warp has lanes
    "allvoid": (1, 3, 21, 5, 5, 4, "low"),         # nothing reaches the 0.3 threshold
}


def make_logits(v, q, c, h, w, seed, kind):
    """Logits shaped like rendered class-probability x mask-probability products: values in [0, 1], a few confident regions."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(v, q, c, h, w, generator=g) * 0.25
    if kind == "low":
        return x
    for qi in range(q):                              # each query: one class, one rectangle per view, confidence above the threshold
        cls = int(torch.randint(0, c, (1,), generator=g))
        if qi < 2:
            cls = qi                                 # make sure both stuff classes (0, 1) occur
        for vi in range(v):
            y0, x0 = int(torch.randint(0, h - 2, (1,), generator=g)), int(torch.randint(0, w - 2, (1,), generator=g))
            y1, x1 = y0 + int(torch.randint(2, h, (1,), generator=g)), x0 + int(torch.randint(2, w, (1,), generator=g))
            x[vi, qi, cls, y0:y1, x0:x1] += 0.3 + 0.6 * torch.rand(1, generator=g)
    if kind == "quantised":
        x = torch.round(x * 4) / 4
    return x


def reference_statements() -> str:
    lines = open(REF).read().splitlines()
    a = next(i for i, l in enumerate(lines) if START in l)
    b = next(i for i, l in enumerate(lines) if STOP in l and i > a)
    return textwrap.dedent("\n".join(lines[a:b]))


def run_reference(logits: torch.Tensor, scores, fuse=(0, 1), num_queries=100):
    ns = {
        "torch": torch,
        "render_output": {"render_qc_logits": [logits.clone()]},
        "context_seg_query_scores": [scores],
        "self": SimpleNamespace(device=torch.device("cpu"),
                                pipecfg=SimpleNamespace(model=SimpleNamespace(mask2former=SimpleNamespace(label_ids_to_fuse=list(fuse), num_queries=num_queries)))),
    }
    exec(compile(reference_statements(), REF, "exec"), ns)
    infos = [{"id": int(i["id"]), "label_id": int(i["label_id"]), "was_fused": bool(i["was_fused"]), "score": float(i["score"])} for i in ns["seg_infos"][0]]
    return ns["all_sem_id"][0].numpy(), ns["all_ins_id"][0].numpy(), infos


def main():
    out, meta = {}, {}
    for name, (v, q, c, h, w, seed, kind) in CASES.items():
        logits = make_logits(v, q, c, h, w, seed, kind)
        scores = [0.5 + 0.05 * i for i in range(q)]
        sem, ins, infos = run_reference(logits, scores)
        out[name + "__logits"] = logits.numpy()
        out[name + "__sem"] = sem.astype(np.int64)
        out[name + "__ins"] = ins.astype(np.int64)
        meta[name] = {"scores": scores, "infos": infos, "shape": [v, q, c, h, w], "label_ids_to_fuse": [0, 1], "num_queries": 100}
        print(name, "sem ids", np.unique(sem).tolist(), "ins ids", np.unique(ins).tolist(), "infos", len(infos))
    out["meta"] = np.array(json.dumps(meta))
    path = os.path.join(ROOT, "tests", "golden", "labels2d_cases.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    sys.exit(main())
