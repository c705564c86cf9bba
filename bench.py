#!/usr/bin/env python
"""bench.py -- headline benchmark of the SIU3R hot path on B200 (contract: see the task statement / DESIGN.md section 6).

  python bench.py --gpus N --steps K --warmup W            our arm (one process per GPU; torchrun for N > 1)
  python bench.py --impl reference --gpus N ...            reference arm: the CPU implementation of the path on the host cores

A "step" = one forward of the per-pair hot path (SIU3RModel.forward) over one batch of synthetic
512x512 image pairs per GPU (BASELINE.json configs[1]: two-view 512x512 inference -> Gaussians + panoptic).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "image_pairs_per_s_512x512"
NSLOTS = int(os.environ.get("SIU3R_BENCH_SLOTS", "2"))   # graph slots alternated by the timed loop
FLOPS_PER_SAMPLE_512_V4 = 8.657e12  # same source: one 4-view 512^2 sample through SIU3RMultiViewModel
FLOPS_PER_PAIR_512 = 4.059e12  # algorithmic 2*MAC FLOPs of one 512^2 pair (FlopCounterMode on the reference)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--batch", type=int, default=1, help="image pairs per GPU per step")
    ap.add_argument("--precision", default="h3", choices=["h3", "tf32", "fp32x3"],
                    help="h3 (default, the mode that meets the north-star tolerances: fp16 hi/lo plane pairs on the kind::f16 tensor-core path), "
                         "tf32 (the reference's own GPU numerics), fp32x3 (round-1 3xTF32 mode)")
    ap.add_argument("--views", type=int, default=2, help="2 = SIU3RModel (headline, BASELINE configs[1]); > 2 = SIU3RMultiViewModel (configs[3])")
    ap.add_argument("--no-multiview", action="store_true", help="skip the short 4-view (configs[3]) measurement appended at N = 1")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-raster", action="store_true")
    ap.add_argument("--graph", type=int, default=1, help="replay the device part of the forward from a CUDA graph")
    return ap.parse_args()


class ClockSampler:
    """Samples nvidia-smi SM clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


class AuxWatchdog:
    """The headline numbers (value, e2e, roofline) are measured first; the sections after them (all-gather, 4-view model, 3xTF32 mode, rasterizer
    sample, CPU baseline) are auxiliary.  If they do not finish within `seconds` (a wedged GPU kernel cannot be interrupted from Python), this
    thread prints the line as it stands -- marked with "aux_timeout" -- and ends the process, so the measured headline is never lost."""

    def __init__(self, line: dict, seconds: float, enabled: bool = True, key: str = "aux_timeout",
                 note: str = "auxiliary sections exceeded {s:.0f} s; printed without the missing ones", exit_code: int = 0):
        self.line, self.seconds, self._stop = line, seconds, threading.Event()
        self.key, self.note, self.exit_code = key, note, exit_code
        if enabled:
            threading.Thread(target=self._run, daemon=True).start()

    def _run(self):
        if not self._stop.wait(self.seconds):
            self.line[self.key] = self.note.format(s=self.seconds)
            try:
                if self.exit_code:   # a wedged headline: leave the Python stacks of all threads on stderr for the post-mortem
                    import faulthandler
                    faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
                print(json.dumps(self.line), flush=True)
            finally:
                os._exit(self.exit_code)

    def cancel(self):
        self._stop.set()


def workload_name(size: int, batch: int = 1, views: int = 2) -> str:
    """config.workload, identical in both arms (the driver compares the strings)."""
    if views == 2:
        return f"two-view {size}x{size} inference -> Gaussians + panoptic (SIU3RModel.forward, enable_query_class_logit_lift=True), {batch} pair(s) per GPU per step"
    return f"{views}-view {size}x{size} inference -> Gaussians + panoptic (SIU3RMultiViewModel.forward, enable_query_class_logit_lift=True), {batch} sample(s) per GPU per step"


def host_threads() -> int:
    """Usable host threads: min(affinity mask, cgroup CPU quota) -- os.cpu_count() over-reports inside containers."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = min(n, max(1, int(float(q) / float(p))))
    except Exception:
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            p = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = min(n, max(1, q // p))
        except Exception:
            pass
    # torch's CPU kernels stop scaling well beyond a few dozen threads on these shapes: probe a GEMM and keep the best
    best, best_t = n, None
    a = torch.randn(2048, 2048)
    for cand in sorted({c for c in (8, 16, 32, 64, n) if c <= n}):
        torch.set_num_threads(cand)
        a @ a
        t0 = time.perf_counter()
        for _ in range(3):
            a @ a
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t * 0.9:
            best, best_t = cand, dt
    return best


def cpu_port_forward(size: int, batch: int, threads: int):
    """One forward of the oracle's PyTorch port on the host cores (the only place bench.py executes oracle/)."""
    from oracle import torch_port as TP
    from siu3r_b200 import synth
    torch.set_num_threads(threads)
    sd = cpu_port_forward.sd if hasattr(cpu_port_forward, "sd") else synth.make_state_dict(populated=True)
    cpu_port_forward.sd = sd
    img, K = synth.pair_inputs(batch, 2, size)
    t0 = time.perf_counter()
    TP.forward(sd, img, K, lift=True)
    return time.perf_counter() - t0


def run_reference(args, rank):
    """Reference arm: the path's CPU implementation (oracle PyTorch port, kind="port": the reference's Python modules live
    under /root/reference and do not exist on the GPU box) on all host cores.  Rank 0 only."""
    if rank != 0:
        return
    threads = host_threads()
    warm = args.warmup                      # ~5 s each at 512^2
    for _ in range(warm):
        cpu_port_forward(args.size, 1, threads)
    ts = [cpu_port_forward(args.size, 1, threads) for _ in range(args.steps)]
    total = sum(ts)
    v = args.steps / total
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": warm,
            "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(args.size, 1, 2),
                       "weights": "seeded random init of the reference architecture (655.5 M params), populated-panoptic preset (siu3r_b200/synth.py)"},
            "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": threads, "kind": "port",
                             "sample": f"{args.steps} x one {args.size}x{args.size} pair, oracle/torch_port.py (torch CPU fp32)"},
            "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    import torch.distributed as dist
    from siu3r_b200 import ops, synth
    from siu3r_b200.model import ModelCfg, SIU3RModel, SIU3RMultiViewModel
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback exists for the product path)"
    torch.cuda.set_device(local)
    from siu3r_b200.parallel import bind_to_gpu_numa
    orig_affinity = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    numa = bind_to_gpu_numa(local) if os.environ.get("SIU3R_NUMA_BIND", "1") != "0" else {"bound": False}
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    S, B = args.size, args.batch
    V = args.views
    model = (SIU3RModel if V == 2 else SIU3RMultiViewModel)(ModelCfg(image_size=(S, S)), precision=args.precision)
    model.load_state_dict(synth.make_state_dict(populated=True))
    model.cuda()
    img_h, K_h = synth.pair_inputs(B, V, S, seed=rank)
    img_pin, K_pin = img_h.pin_memory(), K_h.pin_memory()
    img_d, K_d = img_pin.to(dev, non_blocking=True), K_pin.to(dev, non_blocking=True)
    if args.graph:
        model.enable_cuda_graph()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        barrier()
        ms = s.elapsed_time(e)
        timed.per_rank = [ms]
        if world > 1:
            t = torch.tensor([ms], device=dev)
            allr = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allr, t)
            timed.per_rank = [float(x) for x in allr]
            ms = max(timed.per_rank)
        return ms

    # A GPU kernel that never ends cannot be interrupted from Python: rather than sit until the caller's limit, report where the headline stopped.
    phase = {"metric": METRIC, "value": None, "unit": "pairs/s", "n_gpus": world, "phase": "warm-up / graph capture"}
    guard = AuxWatchdog(phase, float(os.environ.get("SIU3R_BENCH_HANG_TIMEOUT", "300")), enabled=(rank == 0), key="error",
                        note="headline section did not finish within {s:.0f} s (see `phase`); no number was measured", exit_code=3)

    # ---- device-resident throughput (`value`) ----
    # enable_query_class_logit_lift=True as in the reference's inference.py:132-136; with the synthetic weights the panoptic post-process takes its
    # populated branch at 512^2 (19 queries pass the score test, 13 fail the area test, 6 survive, 4 of them fused: siu3r_b200/synth.py)
    step = lambda: model(img_d, K_d, enable_query_class_logit_lift=True)

    def run_steps(n):
        """n forwards of the public API; with the CUDA graph two graph slots alternate, so the device part of step i+1 is already running
        while the host finishes step i (panoptic post-process: a 100-scalar D2H and the label kernels)."""
        if not args.graph:
            for _ in range(n):
                step()
            return
        pend = None
        for i in range(n):
            h = model.forward_async(img_d, K_d, slot=i % NSLOTS)
            if pend is not None:
                model.forward_finish(pend, enable_query_class_logit_lift=True)
            pend = h
        model.forward_finish(pend, enable_query_class_logit_lift=True)

    run_steps(max(args.warmup, 2))
    phase["phase"] = "timed device-resident steps"
    clocks = ClockSampler(local)
    clocks.start()
    ops.reset_launch_count()
    ms = timed(lambda: run_steps(args.steps), 1)
    launches = ops.launch_count()
    clk = clocks.stop()
    value = world * B * args.steps / (ms / 1e3)

    # ---- end to end through the public API with host buffers (`e2e`) ----
    # Every step uploads its pair from pinned host memory and downloads that step's Gaussians + labels into pinned host
    # memory; the download of step i overlaps the forward of step i+1 (siu3r_b200.serving.PairPipeline) and the last
    # one is drained inside the timed region.
    from siu3r_b200.serving import PairPipeline
    phase["phase"] = "end-to-end steps (PairPipeline)"
    pipe = PairPipeline(model, lift=True)

    def e2e_step():
        pipe.submit(img_pin, K_pin)

    def e2e_run(n):
        for _ in range(n):
            e2e_step()
        pipe.flush()
        torch.cuda.current_stream().synchronize()

    e2e_run(3)
    ms_e2e = timed(lambda: e2e_run(args.steps), 1)
    e2e_per_rank = [m / args.steps for m in timed.per_rank]
    e2e_v = world * B * args.steps / (ms_e2e / 1e3)
    host_out = pipe.slots[0]["host"]
    h2d = img_pin.numel() * 4 + K_pin.numel() * 4
    d2h = sum(t.numel() * t.element_size() for t in host_out.values())
    # PCIe download rate of this box (explains e2e vs value: ~200 MB of Gaussians leave the GPU per pair)
    big = max(pipe.slots[0]["dev"].values(), key=lambda t: t.numel())
    big_h = torch.empty(big.shape, dtype=big.dtype, pin_memory=True)
    big_h.copy_(big); torch.cuda.synchronize()
    ms_copy = timed(lambda: big_h.copy_(big, non_blocking=True), 3) / 3
    d2h_gbs = big.numel() * big.element_size() / ms_copy / 1e6

    # ---- roofline of the dominant kernel family (instrumented pass, CUDA events on the launching stream) ----
    phase["phase"] = "instrumented eager pass (roofline)"
    model.disable_cuda_graph()
    model.serial = True   # no parallel branches: per-kernel CUDA-event durations are not inflated by co-running kernels
    step()
    torch.cuda.synchronize()
    # The eager pass issues ~1100 launches at 30-60 us of host time each; a kernel shorter than that would be timed as host latency.  A long
    # spin kernel in front lets the host run ahead, so that every event pair below is stamped by a GPU that never waits for the host.
    torch.cuda._sleep(int(0.45 * 1.9e9))
    ops.PROFILE = []
    step()
    torch.cuda.synchronize()
    model.serial = False
    fam = {}
    shapes = {}
    for f, work, s, e, tag in ops.PROFILE:
        d = fam.setdefault(f, [0.0, 0.0, 0])
        t_ = s.elapsed_time(e)
        d[0] += t_; d[1] += work; d[2] += 1
        if tag is not None:
            sh = shapes.setdefault(tag, [0.0, 0.0, 0])
            sh[0] += t_; sh[1] += work; sh[2] += 1
    if os.environ.get("SIU3R_BENCH_SHAPES") and rank == 0:   # per-shape table of the tensor-core launches (stderr)
        for tag, (t_, wk, n_) in sorted(shapes.items(), key=lambda kv: -kv[1][0])[:40]:
            print(f"{str(tag):44s} x{n_:4d} {t_:8.3f} ms  {wk / (t_ / 1e3) / 1e12:7.1f} TFLOP/s  {1e3 * t_ / n_:7.1f} us/launch", file=sys.stderr)
    ops.PROFILE = None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    bf16_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    h3 = args.precision == "h3"
    if h3:
        # h3 evaluates every algorithmic MAC with three kind::f16 MMAs (hi.hi, lo.hi, hi.lo): the tensor-pipe bound on ALGORITHMIC flop/s is a third of
        # the measured fp16/bf16 rate
        peak_src = ("MEASURED_PEAKS.json bf16_tflops_sustained / 3 (h3: three kind::f16 MMAs per algorithmic MAC)" if peaks
                    else "fallback 1.4 PF/s bf16 sustained / 3")
        tf32_peak = bf16_peak / 3
    else:
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained / 2 (TF32 rate = half the bf16 rate)" if peaks else "fallback 1.4 PF/s bf16 sustained / 2"
        tf32_peak = bf16_peak / 2
    dom = max(fam.items(), key=lambda kv: kv[1][0]) if fam else None
    roofline = None
    # DRAM traffic per launch of the dominant family, from the committed ncu pass over the same forward (profiles/r01_ncu_traffic.json,
    # made by tools/summarize_launches.py --traffic from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`); null if absent
    traffic = {}
    for cand in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", cand)))
            break
        except Exception:
            pass
    if dom:
        name, (tms, work, n) = dom
        ach = work / (tms / 1e3) / 1e12
        tr = traffic.get(name, {}).get("dram_bytes_per_launch") if S == 512 and B == 1 and V == 2 else None
        roofline = {"kernel": name, "bound": "tensor", "achieved": ach, "peak": tf32_peak, "unit": "TFLOP/s", "frac": ach / tf32_peak,
                    "frac_of_tf32_peak": ach / (bf16_peak / 2), "traffic": tr, "traffic_unit": "bytes of DRAM read+write per launch (ncu)", "algorithmic_flop_per_launch": work / n, "launches": n, "ms_per_launch": tms / n, "share_of_step_ms": tms, "peak_source": peak_src,
                    "families": {k: {"ms": v[0], "tflops": v[1] / (v[0] / 1e3) / 1e12, "launches": v[2]} for k, v in fam.items()}}

    line = {"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"tf32": "tf32", "fp32x3": "3xtf32", "h3": "f32 (fp16 hi+lo operand pairs, fp32 accumulate)"}[args.precision], "data": "synthetic",
            "config": {"workload": workload_name(S, B, V), "precision": args.precision,
                       "weights": "seeded random init of the reference architecture (655.5 M params), populated-panoptic preset (siu3r_b200/synth.py)", "parallelism": f"dp{world}",
                       "cuda_graph": bool(args.graph), "overlap": "two graph slots: the device part of step i+1 runs while step i is post-processed" if args.graph else "none",
                       "l2": "no explicit flush: weights (2.6 GB) + activations per step exceed the 126 MB L2 many times over"},
            "clocks": clk,
            "e2e": {"value": e2e_v, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps,
                    "d2h_pinned_GBs": d2h_gbs, "ms_per_step_per_rank": e2e_per_rank, "numa": numa},
            "gpu_launches": launches,
            "vit_tensor_pipe_frac": value / world * (FLOPS_PER_PAIR_512 if V == 2 else FLOPS_PER_SAMPLE_512_V4 * V / 4) * (S / 512.0) ** 2 / 1e12 / tf32_peak,
            "roofline": roofline}
    guard.cancel()
    watchdog = AuxWatchdog(line, float(os.environ.get("SIU3R_BENCH_AUX_TIMEOUT", "300")), enabled=(rank == 0))

    # ---- BASELINE configs[2]: the one collective of the path -- all-gather of the packed render records (85 fp32 per Gaussian) so that every
    #      rank holds the Gaussians of all pairs for joint-scene rasterisation (N > 1 only; not part of `value`) ----
    if world > 1:
        try:
            from siu3r_b200 import parallel
            g0 = model(img_d, K_d)[0]
            rec = parallel.pack_render_record(g0)
            for _ in range(2):
                parallel.all_gather_gaussians(rec)
            ms_ag = timed(lambda: parallel.all_gather_gaussians(rec), 5) / 5
            nbytes = rec.numel() * 4
            line["allgather"] = {"what": "NCCL all_gather_into_tensor of the packed render records (means 3 + cov6 + SH 75 + opacity 1 = 85 fp32 per Gaussian, packed by siu3r_render_record_pack)",
                                 "bytes_contributed_per_rank": nbytes, "bytes_gathered_per_rank": nbytes * world, "ms": ms_ag,
                                 "algbw_GBs": nbytes * world / ms_ag / 1e6, "busbw_GBs": nbytes * (world - 1) / ms_ag / 1e6}
            del g0, rec
        except Exception as ex:
            line["allgather"] = {"error": repr(ex)}

    # ---- BASELINE configs[2] as specified: 4 pairs per GPU per step (batch 32 on 8 GPUs), render records packed on the device and all-gathered INSIDE the
    #      timed step (at N = 1 the same step without the collective) ----
    if V == 2 and B == 1 and not args.no_multiview:
        try:
            from siu3r_b200 import parallel
            model._use_graph = bool(args.graph)     # (the roofline pass above ran eagerly)
            B3 = 4
            i3, K3 = synth.pair_inputs(B3, 2, S, seed=100 + rank)
            i3, K3 = i3.to(dev), K3.to(dev)

            def finish3(h):
                g3 = model.forward_finish(h, enable_query_class_logit_lift=True)[0]
                rec3 = parallel.pack_render_record(g3)
                return parallel.all_gather_gaussians(rec3) if world > 1 else rec3

            def run3(n):   # two graph slots like the headline: post-process, packing and the all-gather of step i run next to the forward of step i + 1
                pend = None
                for i in range(n):
                    h = model.forward_async(i3, K3, slot=i % NSLOTS)
                    if pend is not None:
                        finish3(pend)
                    pend = h
                finish3(pend)
            run3(3)
            n3 = 6
            ms3 = timed(lambda: run3(n3), 1) / n3
            rec_bytes = B3 * 2 * S * S * parallel.RECORD_FLOATS * 4
            line["config3"] = {"workload": f"{B3} pairs per GPU per step ({B3 * world} pairs per step over {world} GPU(s)), forward + device-side record packing"
                                           + (" + NCCL all-gather of the render records, all inside the timed region" if world > 1 else "")
                                           + "; two graph slots as in the headline",
                               "value": world * B3 * 1e3 / ms3, "unit": "pairs/s", "ms_per_step": ms3, "steps": n3,
                               "record_bytes_contributed_per_rank": rec_bytes, "record_bytes_gathered_per_rank": rec_bytes * world}
            del i3, K3
            model._graphs = {k: v for k, v in model._graphs.items() if k[0] == B}   # drop the batch-4 graph (its static buffers) again
            torch.cuda.empty_cache()
        except Exception as ex:
            line["config3"] = {"error": repr(ex)}

    # ---- BASELINE configs[3]: 4-view sample through SIU3RMultiViewModel (short, N = 1 only; not the headline) ----
    if V == 2 and world == 1 and not args.no_multiview:
        try:
            del pipe
            model._graphs = {}
            mv = SIU3RMultiViewModel(ModelCfg(image_size=(S, S)), precision=args.precision)
            mv.w, mv.dev, mv._ready = model.w, model.dev, True      # same packed weights (identical state_dict key set)
            mv.enable_cuda_graph()
            i4, K4 = synth.pair_inputs(1, 4, S, seed=rank)
            i4, K4 = i4.to(dev), K4.to(dev)
            for _ in range(2):
                mv(i4, K4)
            ms4 = timed(lambda: mv(i4, K4), 5) / 5
            line["multiview"] = {"workload": f"4-view {S}x{S} sample (SIU3RMultiViewModel.forward), 1 per step", "samples_per_s": 1e3 / ms4, "ms_per_step": ms4,
                                 "steps": 5, "tensor_pipe_frac": 1e3 / ms4 * FLOPS_PER_SAMPLE_512_V4 * (S / 512.0) ** 2 / 1e12 / tf32_peak}
            del mv
        except Exception as ex:  # never lose the headline line
            line["multiview"] = {"error": repr(ex)}

    # ---- the parity-grade precision mode (3xTF32: north-star tolerances, tests/test_model_gpu.py) timed on the same workload ----
    if V == 2 and world == 1 and not args.no_multiview and args.precision == "h3":
        try:
            del model
            torch.cuda.empty_cache()
            m3 = SIU3RModel(ModelCfg(image_size=(S, S)), precision="tf32")
            m3.load_state_dict(synth.make_state_dict(populated=True))
            m3.cuda()
            m3.enable_cuda_graph()
            for _ in range(2):
                m3(img_d, K_d, enable_query_class_logit_lift=True)
            ms3 = timed(lambda: m3(img_d, K_d, enable_query_class_logit_lift=True), 5) / 5
            line["tf32_mode"] = {"value": B * 1e3 / ms3, "unit": "pairs/s", "ms_per_step": ms3, "steps": 5,
                                 "note": "single-pass TF32 (the reference's own GPU numerics, croco/croco.py:13): outside the north-star tolerances "
                                         "(~2e-3), inside the measured envelope of the reference's TF32 path (profiles/r02_tf32_envelope_S512.json); no slot overlap"}
            del m3
            torch.cuda.empty_cache()
        except Exception as ex:
            line["tf32_mode"] = {"error": repr(ex)}

    # ---- rasterizer sample (BASELINE config 5: 500k pixel-aligned Gaussians @512^2), HBM roofline ----
    if not args.no_raster and rank == 0:
        try:
            line["raster"] = raster_bench(dev, peaks)
        except Exception as ex:  # never lose the headline line
            line["raster"] = {"error": repr(ex)}
    # ---- 2-D labels from rendered query-class logits (30 surviving queries x 21 classes, two 512^2 target views), HBM roofline ----
    if not args.no_raster and rank == 0:
        try:
            line["labels2d"] = labels2d_bench(dev, peaks)
        except Exception as ex:
            line["labels2d"] = {"error": repr(ex)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if orig_affinity is not None:
            os.sched_setaffinity(0, orig_affinity)   # the CPU leg may use every core the process was given, not only the GPU's NUMA node
        threads = host_threads()
        t = cpu_port_forward(S, 1, threads)
        line["cpu_baseline"] = {"value": 1.0 / t, "unit": "pairs/s", "cores": threads, "kind": "port",
                                "sample": f"1 x one {S}x{S} pair (no warm-up), oracle/torch_port.py (torch CPU fp32, restatement pinned to reference goldens)"}
    watchdog.cancel()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def raster_bench(dev, peaks):
    from siu3r_b200 import ops, synth
    from siu3r_b200.renderer import camera_matrices
    G, H, W = 500000, 512, 512
    sc = synth.raster_scene(G, H, W, seed=0, pixel_aligned=True)
    view, full, campos, tx, ty = camera_matrices(sc["extrinsics"], sc["intrinsics"], sc["near"], sc["far"])
    a = [sc[k].to(dev) for k in ("means", "covariances", "harmonics", "opacities")]
    cam = [view[0].to(dev), full[0].to(dev), campos[0].to(dev), torch.zeros(3, device=dev)]
    def fn(touched=True):
        return ops.raster_forward(a[0], a[1], a[2], a[3], cam[0], cam[1], cam[2], cam[3], float(tx[0]), float(ty[0]), H, W, 4, sh_layout=1,
                                  count_touched=touched)
    r = fn()
    D = r["num_rendered"]
    n = 20

    def run(touched):
        for _ in range(3):
            fn(touched)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            fn(touched)
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / n
    ms5 = run(True)         # full 5-tuple of the reference rasterizer (image, radii, depth, opacity, n_touched), host sync per frame to return the duplicate count
    # what render_cuda runs: colour + depth, no n_touched, NO host round trip (siu3r_raster_forward_nosync: status word checked after the batch)
    status = torch.zeros(4, device=dev, dtype=torch.int32)
    st = {"ws": None}

    def fn_ns():
        r_ = ops.raster_forward_nosync(a[0], a[1], a[2], a[3], cam[0], cam[1], cam[2], cam[3], float(tx[0]), float(ty[0]), H, W, 4, sh_layout=1,
                                       status=status, ws=st["ws"], out=st.get("out"))
        st["ws"], st["out"] = r_["ws"], {k: r_[k] for k in ("color", "depth", "opacity", "radii", "n_touched")}
    fn = lambda touched=False: fn_ns()
    ms = run(False)
    assert status.cpu().tolist()[2] == 0
    algo_bytes = 388.0 * G + 68.0 * D + 20.0 * H * W  # preprocess + binning/blend + output
    hbm = peaks.get("hbm_gbs", 6650.0)
    traffic = None
    for cand in ("r02_ncu_raster_traffic.json",):
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", cand))).get("dram_bytes_per_frame")
        except Exception:
            pass
    return {"workload": f"{G} pixel-aligned Gaussians @ {H}x{W}, 1 camera (render_cuda path: colour + depth, no host synchronisation)", "fps": 1e3 / ms, "ms": ms,
            "fps_5tuple_sync": 1e3 / ms5, "ms_5tuple_sync": ms5, "duplicates": D,
            "roofline": {"bound": "hbm", "achieved": algo_bytes / (ms / 1e3) / 1e9, "peak": hbm, "unit": "GB/s", "frac": algo_bytes / (ms / 1e3) / 1e9 / hbm,
                         "traffic": traffic, "traffic_unit": "DRAM bytes read+write per frame, all kernels of the frame (ncu)", "algorithmic_bytes": algo_bytes}}


def labels2d_bench(dev, peaks):
    from siu3r_b200.labels2d import labels_from_qc_logits
    v, q, c, h, w = 2, 30, 21, 512, 512
    x = torch.rand(v, h, w, q * c, device=dev).view(v, h, w, q, c).permute(0, 3, 4, 1, 2)      # the rasteriser's channel-last buffer as "n q c h w"
    scores = [[0.9] * q]
    for _ in range(3):
        labels_from_qc_logits([x], scores)
    torch.cuda.synchronize()
    n = 10
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        labels_from_qc_logits([x], scores)      # includes the q-int download that seg_infos needs
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / n
    algo_bytes = 4.0 * v * q * c * h * w + 16.0 * v * h * w
    hbm = peaks.get("hbm_gbs", 6650.0)
    return {"workload": f"{v} views {h}x{w}, {q} queries x {c} classes (1.32 GB of logits, larger than L2)", "ms": ms, "views_per_s": v * 1e3 / ms,
            "roofline": {"bound": "hbm", "achieved": algo_bytes / (ms / 1e3) / 1e9, "peak": hbm, "unit": "GB/s", "frac": algo_bytes / (ms / 1e3) / 1e9 / hbm,
                         "traffic": None, "algorithmic_bytes": algo_bytes}}


if __name__ == "__main__":
    main()
