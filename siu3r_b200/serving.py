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
    """submit(images, intrinsics) -> finished result of the pair submitted two calls earlier (or None); flush() -> the remaining results.
    Pair n computes (graph slot n % 2) while pair n-1 is post-processed / snapshotted and pair n-2 is being downloaded.

    A result is (host_gaussians: dict[str, pinned tensor], seg_masks, seg_infos).  The pinned tensors of a slot are
    reused `depth` submits later: consume (or copy) them before that.
    """

    def __init__(self, model, depth: int = 3, fields=GAUSSIAN_FIELDS, lift: bool = False):
        assert depth >= 3, "one slot downloading, one being post-processed, one being overwritten"
        self.model, self.depth, self.fields = model, depth, tuple(fields)
        self.lift = lift   # enable_query_class_logit_lift (inference.py:132-136); the lifted logits stay on the device like in the reference
        #                    (Gaussians.detach_cpu_copy only moves tensor attributes; seg_query_class_logits is a list)
        self.copy_stream = torch.cuda.Stream(device=model.dev)
        self.slots = [dict(dev={}, host={}, snap=torch.cuda.Event(), done=torch.cuda.Event(), meta=None) for _ in range(depth)]
        self.n = 0
        self._pending = None
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _collect(self, slot):
        slot["done"].synchronize()
        return slot["host"], *slot["meta"]

    def submit(self, images: torch.Tensor, intrinsics: torch.Tensor):
        """Enqueue pair n; returns the finished result of pair n-1 (None for the first call).  The forward of pair n (graph slot n % 2) is
        already running on the GPU while pair n-1 is post-processed, snapshotted and downloaded."""
        if images.device.type == "cpu":   # forward_async() uploads (pinned -> device, asynchronous) itself
            self.h2d_bytes = images.numel() * images.element_size() + intrinsics.numel() * intrinsics.element_size()
        handle = self.model.forward_async(images, intrinsics, slot=self.n % 2)
        prev = self._retire()
        self._pending = (self.n, handle)
        self.n += 1
        return prev

    def _retire(self):
        """Finish the pending forward: post-process, device-side snapshot, asynchronous download; returns the PREVIOUS finished result."""
        if self._pending is None:
            return None
        idx, handle = self._pending
        self._pending = None
        cur = torch.cuda.current_stream()
        slot = self.slots[idx % self.depth]
        out = self.model.forward_finish(handle, enable_query_class_logit_lift=self.lift)
        g, seg_masks, seg_infos = out[0], out[2], out[3]
        if idx >= self.depth:
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
        prev_idx = idx - 1
        ready = self._collect(self.slots[prev_idx % self.depth]) if prev_idx >= 0 else None
        self._last = idx
        return ready

    def flush(self) -> list:
        """Drain: returns the results that submit() has not handed out yet (at most two), in submission order."""
        if self.n == 0:
            return []
        out = []
        if self._pending is not None:
            r = self._retire()
            if r is not None:
                out.append(r)
        out.append(self._collect(self.slots[(self.n - 1) % self.depth]))
        return out
