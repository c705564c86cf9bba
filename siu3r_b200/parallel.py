"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch; gloo in CPU tests).

The per-pair path shards with NO data-path collective: pairs are independent units (every norm layer is
per-sample in eval mode).  The single collective of the north star is the all-gather of the packed per-Gaussian render
record, issued only when Gaussians of several pairs must be merged into one scene for joint rasterisation.
There is no precedent in the reference (its inference is single-GPU: inference.py:116-124).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

RECORD_FLOATS = 3 + 6 + 75 + 1  # means, covariance upper triangle (what the rasterizer consumes), harmonics (3x25), opacity


def bind_to_gpu_numa(device_index: int) -> dict:
    """Pin the calling process to the CPUs next to GPU `device_index` BEFORE it allocates pinned host buffers, so that the ~200 MB of Gaussians every pair
    downloads (Gaussians.detach_cpu_copy, gaussians_types.py:25-38) cross the GPU's own PCIe root instead of the socket interconnect.  One process per
    GPU (torchrun) is assumed; a restriction that would leave no allowed CPU is skipped.  Returns what was done (for the bench line)."""
    import os
    info = {"device": device_index, "bound": False}
    try:
        bus = torch.cuda.get_device_properties(device_index)
        pci = f"{bus.pci_domain_id:04x}:{bus.pci_bus_id:02x}:{bus.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{pci}"
        node = int(open(base + "/numa_node").read())
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            if part:
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        info.update(numa_node=node, local_cpus=len(cpus), allowed_cpus=len(allowed))
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            info["bound"] = True
            info["cpus"] = len(use)
    except Exception as e:   # no sysfs / no permission: keep the inherited affinity
        info["error"] = repr(e)[:120]
    return info


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Contiguous block partition: rank r owns items [r*ceil.., ...) (batch 32 on 8 GPUs -> pairs 4r .. 4r+3)."""
    per = (n_items + world - 1) // world
    lo = min(n_items, rank * per)
    return range(lo, min(n_items, lo + per))


def pack_render_record(g) -> torch.Tensor:
    """Gaussians (batch b) -> [b, G, 85] fp32 render record.  CUDA tensors go through siu3r_render_record_pack (one coalesced pass, no torch
    compute op); the CPU form below only serves the gloo tests of the host logic."""
    b, G = g.means.shape[:2]
    if g.means.is_cuda:
        from . import _lib, ops
        out = torch.empty(b, G, RECORD_FLOATS, device=g.means.device, dtype=torch.float32)
        m, c, h, o = g.means.contiguous(), g.covariances.contiguous(), g.harmonics.contiguous(), g.opacities.contiguous()
        _lib.check(_lib.load().siu3r_render_record_pack(m.data_ptr(), c.data_ptr(), h.data_ptr(), o.data_ptr(), b * G, out.data_ptr(), ops._stream()),
                   "render_record_pack")
        return out
    row, col = torch.triu_indices(3, 3)
    return torch.cat([g.means.reshape(b, G, 3), g.covariances[:, :, row, col], g.harmonics.reshape(b, G, 75), g.opacities.reshape(b, G, 1)], dim=-1).contiguous()


def unpack_render_record(rec: torch.Tensor):
    """[b, G, 85] -> means [b,G,3], cov6 [b,G,6] (rasterizer cov_stride = 6), harmonics [b,G,3,25], opacities [b,G]."""
    b, G = rec.shape[:2]
    if rec.is_cuda:
        from . import _lib, ops
        dev = rec.device
        means, cov6 = torch.empty(b, G, 3, device=dev), torch.empty(b, G, 6, device=dev)
        harm, opac = torch.empty(b, G, 3, 25, device=dev), torch.empty(b, G, device=dev)
        _lib.check(_lib.load().siu3r_render_record_unpack(rec.contiguous().data_ptr(), b * G, means.data_ptr(), cov6.data_ptr(), harm.data_ptr(),
                                                          opac.data_ptr(), ops._stream()), "render_record_unpack")
        return means, cov6, harm, opac
    return (rec[..., 0:3].contiguous(), rec[..., 3:9].contiguous(), rec[..., 9:84].reshape(b, G, 3, 25).contiguous(), rec[..., 84].contiguous())


def all_gather_gaussians(rec: torch.Tensor, group=None) -> torch.Tensor:
    """ONE all-gather of the packed records: [b_local, G, 85] per rank -> [world * b_local, G, 85] on every rank."""
    world = dist.get_world_size(group)
    out = torch.empty((world * rec.shape[0],) + tuple(rec.shape[1:]), dtype=rec.dtype, device=rec.device)
    dist.all_gather_into_tensor(out, rec.contiguous(), group=group)
    return out
