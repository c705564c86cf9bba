"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path (siu3r_b200/).

NumPy restatement of the 2-D label extraction the reference applies to the rendered query-class logits in its validation / test
step (/root/reference/src/pipeline.py:132-193; the viewer repeats it, viewer.py:404-446).  Pinned: tests/golden/labels2d_cases.npz
holds outputs of the reference's own statements (executed unmodified from /root/reference by oracle/make_golden_labels2d.py) and
tests/test_oracle_cpu.py checks this restatement against them.

For one sample, logits [v, q, c, h, w] (c = classes + void, void LAST), q_scores [q]:
    :143      c_logit, q_index = max over q                                  -> [v, c, h, w]
    :145-150  move the void channel to the front of the class axis
    :151      sem_logit, sem_id = max over the rotated class axis            -> [v, h, w]
    :152-161  q_index = q_index[.., sem_id, ..] + 1
    :162-164  sem_id = 0 where sem_logit < 0.3;  q_index = 0 where sem_id == 0
    :165-180  seg info per query that owns at least one pixel: label = sem_id of its first pixel in (v, h, w) order
    :182-191  stuff classes: instance id num_queries + stuff + 1 on their pixels; their infos get was_fused and that id
"""
from __future__ import annotations

import numpy as np


def labels_from_qc_logits(logits: np.ndarray, q_scores, label_ids_to_fuse=(0, 1), num_queries: int = 100, threshold: float = 0.3):
    """-> (sem_id [v,h,w] int64, ins_id [v,h,w] int64, infos list of dict(id, label_id, was_fused, score))."""
    logits = np.asarray(logits, np.float32)
    v, q, c, h, w = logits.shape
    c_logit = logits.max(axis=1)                       # [v, c, h, w]
    q_index = logits.argmax(axis=1)                    # first maximum, like torch.max on the CPU
    order = [c - 1] + list(range(c - 1))
    c_logit, q_index = c_logit[:, order], q_index[:, order]
    sem_logit = c_logit.max(axis=1)
    sem_id = c_logit.argmax(axis=1).astype(np.int64)
    ins_id = np.take_along_axis(q_index, sem_id[:, None], axis=1)[:, 0].astype(np.int64) + 1
    sem_id[sem_logit < np.float32(threshold)] = 0
    ins_id[sem_id == 0] = 0
    infos = []
    flat_sem, flat_ins = sem_id.reshape(-1), ins_id.reshape(-1)
    for qi, score in enumerate(q_scores):
        own = np.flatnonzero(flat_ins == qi + 1)
        if own.size == 0:
            continue
        infos.append({"id": qi + 1, "label_id": int(flat_sem[own[0]]), "was_fused": False, "score": score})
    for stuff in label_ids_to_fuse:
        m = sem_id == stuff + 1
        ins_id[m] = num_queries + stuff + 1
        for i in infos:
            if i["label_id"] == stuff + 1:
                i["was_fused"] = True
                i["id"] = int(ins_id[m][0])
    return sem_id, ins_id, infos
