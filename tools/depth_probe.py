"""Throughput of the headline workload with 1, 2, 3 and 4 forwards in flight (graph slots): does more overlap fill the machine further?"""
import os, sys, time, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from siu3r_b200 import synth
from siu3r_b200.model import ModelCfg, SIU3RModel
S = 512
model = SIU3RModel(ModelCfg(image_size=(S, S)), precision="h3")
model.load_state_dict(synth.make_state_dict(populated=True))
model.cuda()
model.enable_cuda_graph()
img, K = synth.pair_inputs(1, 2, S)
img, K = img.cuda(), K.cuda()


def run(n, depth):
    pend = []
    for i in range(n):
        pend.append(model.forward_async(img, K, slot=i % depth))
        if len(pend) >= depth:
            model.forward_finish(pend.pop(0), enable_query_class_logit_lift=True)
    while pend:
        model.forward_finish(pend.pop(0), enable_query_class_logit_lift=True)


out = {}
for depth in (1, 2, 3, 4):
    run(2 * depth + 2, depth)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(24, depth)
    torch.cuda.synchronize()
    out[f"depth{depth}_ms"] = (time.perf_counter() - t0) / 24 * 1e3
print(json.dumps(out))
