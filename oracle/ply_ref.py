"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path (siu3r_b200/).

NumPy restatement of /root/reference/src/utils/ply_export.py:30-97 (export_ply): same attribute order, same float64 concatenate
followed by the per-field cast into the structured vertex dtype, and the header plyfile 1.x writes for
PlyData([PlyElement.describe(elements, "vertex")]) (binary_little_endian, 'float' / 'int' property names, no comments).
PINNED for the record layout: tests/golden/ply_records.npz holds the structured records the reference's own export_ply assembles (run unmodified by
oracle/make_golden_ply.py with `plyfile` replaced by a recorder) and tests/test_oracle_cpu.py checks this restatement against them, field by
field (log(scales) to 1 ulp: numpy vs ATen logf).  Still UNPINNED: the ASCII header, which is plyfile's own text (un-vendored, not installed).
"""
from __future__ import annotations

import numpy as np


def construct_list_of_attributes(num_rest: int) -> list:   # ply_export.py:12-27
    a = ["x", "y", "z", "nx", "ny", "nz"] + [f"f_dc_{i}" for i in range(3)] + [f"f_rest_{i}" for i in range(num_rest)] + ["opacity"]
    a += [f"scale_{i}" for i in range(3)] + [f"rot_{i}" for i in range(4)] + ["semantic_label", "instance_label"]
    return a


def export_ply_bytes(means, scales, rotations, harmonics, opacities, semantic_labels, instance_labels, seg_query_class_logits,
                     save_sh_dc_only=True) -> bytes:
    """All inputs numpy arrays (float32 / int32).  Returns the full file contents."""
    x, y, z, w = rotations.T                                             # :52 "g xyzw -> xyzw g"
    rot = np.stack((w, x, y, z), axis=-1)                                # :53
    f_dc = harmonics[..., 0]                                             # :57
    f_rest = harmonics[..., 1:].reshape(harmonics.shape[0], -1)          # :58
    attrs = construct_list_of_attributes(0 if save_sh_dc_only else f_rest.shape[1])
    dtype_full = [(a, "<f4") for a in attrs[:-2]]                        # :63
    if semantic_labels is not None and instance_labels is not None:
        dtype_full += [("semantic_label", "<i4"), ("instance_label", "<i4")]
    if seg_query_class_logits is not None:
        g, q, c = seg_query_class_logits.shape
        seg_query_class_logits = seg_query_class_logits.reshape(g, q * c)
        dtype_full += [(f"seg_query_class_logits_{i}", "<f4") for i in range(q * c)]
    elements = np.empty(means.shape[0], dtype=dtype_full)
    cols = [means, np.zeros_like(means), f_dc, f_rest, opacities[..., None], np.log(scales), rot]   # :73-81
    if semantic_labels is not None and instance_labels is not None:
        cols += [semantic_labels[..., None], instance_labels[..., None]]
    if seg_query_class_logits is not None:
        cols.append(seg_query_class_logits)
    if save_sh_dc_only:
        cols.pop(3)
    table = np.concatenate(cols, axis=1)                                  # :92 (promotes to float64 when labels are present)
    for i, (name, _) in enumerate(dtype_full):                            # :93 elements[:] = list(map(tuple, attributes)), column-wise
        elements[name] = table[:, i]
    names = {"<f4": "float", "<i4": "int"}
    header = ["ply", "format binary_little_endian 1.0", f"element vertex {means.shape[0]}"]
    header += [f"property {names[t]} {n}" for n, t in dtype_full] + ["end_header"]
    return ("\n".join(header) + "\n").encode("ascii") + elements.tobytes()
