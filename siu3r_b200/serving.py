"""Throughput-oriented driver around SIU3RModel.forward for a stream of image pairs held in HOST memory.

The reference's inference loop (inference.py:119-141) is strictly serial: upload the pair, run the model, download the
Gaussians (`gaussians.detach_cpu_copy()`, src/utils/gaussians_types.py:25-38), next pair.  On a B200 the download of one
pair's Gaussians (about 200 MB at 512x512) costs several milliseconds of PCIe time, comparable to a fifth of the forward
itself, so the engine overlaps it: the results of pair i are snapshotted device-to-device (microseconds of HBM time) and
drained to pinned host memory on a copy stream while pair i+1 is already computing.
"""
from __future__ import annotations

import torch

GAUSSIAN_FIELDS = ("means", "covariances", "harmonics", "opacities", "scales", "rotations", "semantic_labels", "instance_labels")


class PairPipeline:
    """submit(images, intrinsics) -> result of the PREVIOUS submit (or None); flush() -> result of the last one.

    A result is (host_gaussians: dict[str, pinned tensor], seg_masks, seg_infos).  The pinned tensors of a slot are
    reused `depth` submits later: consume (or copy) them before that.
    """

    def __init__(self, model, depth: int = 2, fields=GAUSSIAN_FIELDS):
        assert depth >= 2
        self.model, self.depth, self.fields = model, depth, tuple(fields)
        self.copy_stream = torch.cuda.Stream(device=model.dev)
        self.slots = [dict(dev={}, host={}, snap=torch.cuda.Event(), done=torch.cuda.Event(), meta=None) for _ in range(depth)]
        self.n = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _collect(self, slot):
        slot["done"].synchronize()
        return slot["host"], *slot["meta"]

    def submit(self, images: torch.Tensor, intrinsics: torch.Tensor):
        cur = torch.cuda.current_stream()
        slot = self.slots[self.n % self.depth]
        prev = self.slots[(self.n - 1) % self.depth] if self.n > 0 else None
        if images.device.type == "cpu":   # forward() uploads (pinned -> device, asynchronous) itself
            self.h2d_bytes = images.numel() * images.element_size() + intrinsics.numel() * intrinsics.element_size()
        out = self.model(images, intrinsics)
        g, seg_masks, seg_infos = out[0], out[2], out[3]
        if self.n >= self.depth:
            cur.wait_event(slot["done"])   # the slot's previous download has left the snapshot buffers
        nbytes = 0
        for name in self.fields:
            t = getattr(g, name)
            if name not in slot["dev"]:
                slot["dev"][name] = torch.empty_like(t)
                slot["host"][name] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            slot["dev"][name].copy_(t, non_blocking=True)   # snapshot: the model's (graph-static) outputs are free again
            nbytes += t.numel() * t.element_size()
        self.d2h_bytes = nbytes
        slot["snap"].record(cur)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(slot["snap"])
            for name in self.fields:
                slot["host"][name].copy_(slot["dev"][name], non_blocking=True)
            slot["done"].record(self.copy_stream)
        slot["meta"] = (seg_masks, seg_infos)
        self.n += 1
        return self._collect(prev) if prev is not None else None

    def flush(self):
        if self.n == 0:
            return None
        return self._collect(self.slots[(self.n - 1) % self.depth])
