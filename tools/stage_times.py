"""Stage time-stamps inside the captured forward graph (external event-record nodes): where the critical path is.
usage: python tools/stage_times.py [--size 512] [--precision tf32]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from siu3r_b200 import synth
from siu3r_b200.model import ModelCfg, SIU3RModel

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--precision", default="tf32")
ap.add_argument("--serial", action="store_true")
a = ap.parse_args()
model = SIU3RModel(ModelCfg(image_size=(a.size, a.size)), precision=a.precision)
model.load_state_dict(synth.make_state_dict())
model.cuda()
model.serial = a.serial
img, K = synth.pair_inputs(1, 2, a.size)
img, K = img.cuda(), K.cuda()
model.enable_cuda_graph()
model.marks = {}
acc = {}
for it in range(8):
    model(img, K)
    torch.cuda.synchronize()
    if it >= 3:
        for n, ev in model.marks.items():
            acc.setdefault(n, []).append(model.marks["start"].elapsed_time(ev))
for n, v in sorted(acc.items(), key=lambda kv: sum(kv[1])):
    print(f"{n:10s} t = {sum(v) / len(v):8.3f} ms after start")
