"""Generates tests/golden/ply_records.npz by running the reference's OWN export_ply (/root/reference/src/utils/ply_export.py:30-97) unmodified.

`plyfile` (un-vendored, not installed) is replaced at the import boundary by a recorder: PlyElement.describe(elements, "vertex") receives the
structured numpy array the reference assembled -- field names, dtypes and the record bytes, i.e. everything export_ply itself decides -- and
PlyData(...).write() stores nothing.  What plyfile would add on disk (the ASCII header) remains a restatement (oracle/ply_ref.py, "header parity
unpinned"); the record layout, log(scales), the xyzw -> wxyz rotation order, label columns and the flattened query-class logits are pinned here.

    python oracle/make_golden_ply.py            (needs /root/reference; CPU only)
"""
from __future__ import annotations

import json
import os
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_model import _import_reference  # noqa: E402

CAPTURED = {}


def scene(G=257, d_sh=25, q=3, c=21, seed=3):
    g = torch.Generator().manual_seed(seed)
    return dict(means=torch.randn(G, 3, generator=g), scales=torch.rand(G, 3, generator=g) * 0.3 + 1e-3, rotations=torch.randn(G, 4, generator=g),
                harmonics=torch.randn(G, 3, d_sh, generator=g), opacities=torch.rand(G, generator=g),
                semantic_labels=torch.randint(0, 21, (G,), generator=g, dtype=torch.int32),
                instance_labels=torch.randint(0, 9, (G,), generator=g, dtype=torch.int32),
                seg_query_class_logits=torch.rand(G, q, c, generator=g))


def main():
    _import_reference()
    pf = types.ModuleType("plyfile")

    class PlyElement:
        @staticmethod
        def describe(elements, name):
            CAPTURED["elements"], CAPTURED["name"] = elements.copy(), name
            return (elements, name)

    class PlyData:
        def __init__(self, els):
            self.els = els

        def write(self, path):
            CAPTURED["path"] = str(path)

    pf.PlyElement, pf.PlyData = PlyElement, PlyData
    sys.modules["plyfile"] = pf
    from src.utils.ply_export import export_ply
    out = {}
    meta = {}
    for tag, dc_only, with_qc, with_labels in (("full", False, True, True), ("dc", True, False, True), ("dc_qc", True, True, True), ("nolabels", False, False, False)):
        s = scene()
        export_ply(means=s["means"], scales=s["scales"], rotations=s["rotations"], harmonics=s["harmonics"], opacities=s["opacities"],
                   semantic_labels=s["semantic_labels"] if with_labels else None, instance_labels=s["instance_labels"] if with_labels else None,
                   seg_query_class_logits=s["seg_query_class_logits"] if with_qc else None, path=Path("/tmp/_siu3r_golden.ply"),
                   shift_and_scale=False, save_sh_dc_only=dc_only)
        el = CAPTURED["elements"]
        out[tag + "__records"] = np.frombuffer(el.tobytes(), dtype=np.uint8)
        meta[tag] = dict(names=list(el.dtype.names), formats=[el.dtype[n].str for n in el.dtype.names], count=int(el.shape[0]), element=CAPTURED["name"],
                         dc_only=dc_only, with_qc=with_qc, with_labels=with_labels)
        print(tag, len(el.dtype.names), "fields", el.dtype.itemsize, "bytes per record")
    out["meta"] = np.array(json.dumps(meta))
    path = os.path.join(ROOT, "tests", "golden", "ply_records.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
