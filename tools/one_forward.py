"""One eager forward (+ one raster call) inside a cudaProfilerStart/Stop window, for `ncu --profile-from-start off`."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from siu3r_b200 import ops, synth
from siu3r_b200.model import ModelCfg, SIU3RModel
from siu3r_b200.renderer import camera_matrices
S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
prec = sys.argv[2] if len(sys.argv) > 2 else "tf32"
model = SIU3RModel(ModelCfg(image_size=(S, S)), precision=prec)
model.load_state_dict(synth.make_state_dict())
model.cuda()
img, K = synth.pair_inputs(1, 2, S)
img, K = img.cuda(), K.cuda()
model(img, K)
G, H, W = 500000, 512, 512
sc = synth.raster_scene(G, H, W, seed=0, pixel_aligned=True)
view, full, campos, tx, ty = camera_matrices(sc["extrinsics"], sc["intrinsics"], sc["near"], sc["far"])
a = [sc[k].cuda() for k in ("means", "covariances", "harmonics", "opacities")]
cam = [view[0].cuda(), full[0].cuda(), campos[0].cuda(), torch.zeros(3, device="cuda")]
rf = lambda: ops.raster_forward(a[0], a[1], a[2], a[3], cam[0], cam[1], cam[2], cam[3], float(tx[0]), float(ty[0]), H, W, 4, sh_layout=1)
rf()
torch.cuda.synchronize()
torch.cuda.profiler.start()
model(img, K)
rf()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
