"""A few frames of the bench.py raster workload (500 k pixel-aligned Gaussians @ 512x512) for `ncu -k regex:render_kernel -s 2 -c 1`."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from siu3r_b200 import ops, synth
from siu3r_b200.renderer import camera_matrices
G, H, W = (int(sys.argv[1]) if len(sys.argv) > 1 else 500000), 512, 512
sc = synth.raster_scene(G, H, W, seed=0, pixel_aligned=True)
view, full, campos, tx, ty = camera_matrices(sc["extrinsics"], sc["intrinsics"], sc["near"], sc["far"])
a = [sc[k].cuda() for k in ("means", "covariances", "harmonics", "opacities")]
cam = [view[0].cuda(), full[0].cuda(), campos[0].cuda(), torch.zeros(3, device="cuda")]
status = torch.zeros(4, device="cuda", dtype=torch.int32)
ws = None
for _ in range(4):
    r = ops.raster_forward_nosync(a[0], a[1], a[2], a[3], cam[0], cam[1], cam[2], cam[3], float(tx[0]), float(ty[0]), H, W, 4, sh_layout=1, status=status, ws=ws)
    ws = r["ws"]
torch.cuda.synchronize()
print("done", status.tolist())
