"""BASELINE config 5: 3DGS raster sweep, G in {100k .. 2M} x {512x512, 1920x1080} x {pixel-aligned, adapter-scale} splats on one GPU.
Reports frames/s, duplicates D, algorithmic bytes (388 G + 68 D + 20 HW) and the fraction of the measured HBM peak.
usage: python tools/raster_sweep.py [--out gpurun_out/raster_sweep.json] [--quick]"""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from siu3r_b200 import ops, synth
from siu3r_b200.renderer import camera_matrices

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="gpurun_out/raster_sweep.json")
ap.add_argument("--quick", action="store_true")
a = ap.parse_args()
dev = "cuda"
try:
    hbm = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0)
except Exception:
    hbm = 6650.0
Gs = [100_000, 500_000] if a.quick else [100_000, 200_000, 500_000, 1_000_000, 2_000_000]
rows = []
for (H, W) in ((512, 512), (1080, 1920)):
    for pa in (True, False):
        for G in Gs:
            sc = synth.raster_scene(G, H, W, seed=0, pixel_aligned=pa)
            view, full, campos, tx, ty = camera_matrices(sc["extrinsics"], sc["intrinsics"], sc["near"], sc["far"])
            t = [sc[k].to(dev) for k in ("means", "covariances", "harmonics", "opacities")]
            cam = [view[0].to(dev), full[0].to(dev), campos[0].to(dev), torch.zeros(3, device=dev)]
            res = {}
            for touched in (True, False):
                fn = lambda: ops.raster_forward(t[0], t[1], t[2], t[3], cam[0], cam[1], cam[2], cam[3], float(tx[0]), float(ty[0]), H, W, 4,
                                                sh_layout=1, count_touched=touched)
                r = fn()
                D = r["num_rendered"]
                cap = int(D * 1.05) + 1024
                fn = lambda: ops.raster_forward(t[0], t[1], t[2], t[3], cam[0], cam[1], cam[2], cam[3], float(tx[0]), float(ty[0]), H, W, 4,
                                                sh_layout=1, count_touched=touched, dup_capacity=cap)
                for _ in range(2):
                    fn()
                torch.cuda.synchronize()
                n = 10
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                for _ in range(n):
                    fn()
                e.record()
                torch.cuda.synchronize()
                res[touched] = s.elapsed_time(e) / n
            # the path render_cuda takes since round 2: no host synchronisation, device status words, reused workspace
            status = torch.zeros(4, device=dev, dtype=torch.int32)
            ws = None
            def fn_ns():
                global ws_keep
                return ops.raster_forward_nosync(t[0], t[1], t[2], t[3], cam[0], cam[1], cam[2], cam[3], float(tx[0]), float(ty[0]), H, W, 4, sh_layout=1,
                                                 status=status, ws=ws, dup_capacity=cap)
            r0 = fn_ns(); ws = r0["ws"]
            for _ in range(2):
                fn_ns()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(n):
                fn_ns()
            e.record()
            torch.cuda.synchronize()
            ms_ns = s.elapsed_time(e) / n
            flags = int(status.cpu()[2])
            ab = 388.0 * G + 68.0 * D + 20.0 * H * W
            row = dict(H=H, W=W, G=G, splats="pixel-aligned" if pa else "adapter-scale", duplicates=D, visible=int((r["radii"] > 0).sum()),
                       ms_full_tuple=res[True], fps_full_tuple=1e3 / res[True], ms_sync_no_touched=res[False], ms_render_cuda=ms_ns, fps_render_cuda=1e3 / ms_ns,
                       nosync_flags=flags, algorithmic_MB=ab / 1e6, achieved_GBs=ab / ms_ns / 1e6, hbm_frac=ab / ms_ns / 1e6 / hbm)
            rows.append(row)
            print(json.dumps(row), flush=True)
            del t, sc
            torch.cuda.empty_cache()
os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
json.dump({"hbm_peak_GBs": hbm, "rows": rows}, open(a.out, "w"), indent=1)
