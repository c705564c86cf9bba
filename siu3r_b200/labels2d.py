"""2-D semantic / instance label maps from rendered query-class logits (second half).

Mirrors what the reference does right after `SplattingCUDA.forward(..., render_qc_logits=True)`:
  labels_from_qc_logits   <-> Pipeline.step_w_query_class_logit_lift, src/pipeline.py:132-193 (all_sem_id, all_ins_id, seg_infos)
  viewer_labels           <-> Viewer._qc_logits_render_fn, viewer.py:422-435 (same arithmetic, stuff ids hard-coded)

The per-pixel arithmetic (max over queries, void-first class rotation, max over classes, 0.3 threshold, instance ids, stuff fusing)
and the "first pixel each query owns" search run in ONE pass over the logits on the GPU (csrc/labels2d.cu); the reference makes ~15
full-size temporaries (two concatenations, a 3-D meshgrid gather, boolean masks) and one masked select + .item() per query.
Only the per-query records (q ints) come back to the host to build seg_infos.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib, ops


def _extract(logits: torch.Tensor, fuse_pairs, threshold: float, sem: torch.Tensor = None, ins: torch.Tensor = None):
    """logits [v, q, c, h, w] float32 CUDA tensor with any strides -> (sem_id [v,h,w] int64, ins_id [v,h,w] int64, first_sem [q] int32 device).
    sem / ins: optional contiguous [v, h, w] int64 outputs to fill."""
    assert logits.is_cuda and logits.dtype == torch.float32 and logits.dim() == 5, "render_qc_logits must be a float32 CUDA tensor [v, q, c, h, w]"
    lib = _lib.load()
    v, q, c, h, w = logits.shape
    assert len(fuse_pairs) <= 8
    dev = logits.device
    if sem is None:
        sem = torch.empty(v, h, w, device=dev, dtype=torch.int64)
        ins = torch.empty(v, h, w, device=dev, dtype=torch.int64)
    assert sem.shape == (v, h, w) and ins.shape == (v, h, w) and sem.is_contiguous() and ins.is_contiguous() and sem.dtype == ins.dtype == torch.int64
    work = torch.empty(2, q, device=dev, dtype=torch.int32)
    n = len(fuse_pairs)
    fs = (ctypes.c_int * max(n, 1))(*[int(a) for a, _ in fuse_pairs])
    fi = (ctypes.c_int * max(n, 1))(*[int(b) for _, b in fuse_pairs])
    sv, sq, sc, sh, sw = logits.stride()
    _lib.check(lib.siu3r_labels_from_qc_logits(logits.data_ptr(), v, q, c, h, w, sv, sq, sc, sh, sw, float(threshold),
                                               ctypes.cast(fs, ctypes.c_void_p), ctypes.cast(fi, ctypes.c_void_p), n, sem.data_ptr(), ins.data_ptr(),
                                               work[0].data_ptr(), work[1].data_ptr(), ops._stream()), "labels_from_qc_logits")
    return sem, ins, work[1]


def labels_from_qc_logits(render_qc_logits, context_seg_query_scores, label_ids_to_fuse=(0, 1), num_queries: int = 100,
                          threshold: float = 0.3):
    """render_qc_logits: list over the batch of [v, q, c+1, h, w] tensors (`SplattingCUDA.forward(...)["render_qc_logits"]`);
    context_seg_query_scores: list over the batch of the q scores of the surviving queries (5th output of SIU3RModel.forward).
    -> (all_sem_id [b, v, h, w] int64, all_ins_id [b, v, h, w] int64, seg_infos: list over the batch of
        [{"id", "label_id", "was_fused", "score"}, ...])   -- pipeline.py:133-193."""
    pairs = [(int(s) + 1, int(num_queries) + int(s) + 1) for s in label_ids_to_fuse]
    b = len(render_qc_logits)
    v, _, _, h, w = render_qc_logits[0].shape
    assert all(t.shape[0] == v and t.shape[3:] == (h, w) for t in render_qc_logits), "all samples must share (v, h, w) (torch.stack at pipeline.py:194-195)"
    dev = render_qc_logits[0].device
    all_sem = torch.empty(b, v, h, w, device=dev, dtype=torch.int64)
    all_ins = torch.empty(b, v, h, w, device=dev, dtype=torch.int64)
    firsts = [_extract(logits, pairs, threshold, all_sem[bi], all_ins[bi])[2]      # launches for the whole batch first, downloads afterwards
              for bi, logits in enumerate(render_qc_logits)]
    seg_infos = [seg_infos_from_first_labels(f.cpu().tolist(), q_score, pairs) for f, q_score in zip(firsts, context_seg_query_scores)]
    return all_sem, all_ins, seg_infos


def seg_infos_from_first_labels(first_sem, q_score, fuse_pairs):
    """pipeline.py:165-191 on the host: first_sem[q] = semantic id of the first pixel query q owns (-1: none, the query is dropped, :168-169);
    infos whose label is a fused stuff class take that class's instance id (a stuff label in an info implies that some pixel carries it)."""
    info = []
    for q_idx, q_score_i in enumerate(q_score):
        if first_sem[q_idx] < 0:
            continue
        info.append({"id": q_idx + 1, "label_id": first_sem[q_idx], "was_fused": False, "score": q_score_i})
    for sem_value, ins_value in fuse_pairs:
        for i in info:
            if i["label_id"] == sem_value:
                i["was_fused"] = True
                i["id"] = ins_value
    return info


def viewer_labels(render_qc_logit: torch.Tensor, semantic_threshold: float = 0.3):
    """viewer.py:422-435: logits [v, q, c+1, h, w] -> (sem_id, q_index) [v, h, w] int64 with the viewer's hard-coded stuff ids
    (sem 1 -> 102, sem 2 -> 103; note this differs from the pipeline's 101 / 102)."""
    sem, ins, _ = _extract(render_qc_logit, [(1, 102), (2, 103)], semantic_threshold)
    return sem, ins
