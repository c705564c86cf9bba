"""Where does end-to-end time go at N > 1?  Run under torchrun; every rank prints its own numbers (no collectives inside the timed loops).
usage: python -m torch.distributed.run --nproc-per-node 2 ... tools/e2e_probe.py"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1 and os.environ.get("PROBE_NCCL", "1") == "1":
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dist.barrier()
from siu3r_b200 import synth
from siu3r_b200.model import ModelCfg, SIU3RModel
from siu3r_b200.serving import PairPipeline, GAUSSIAN_FIELDS
S = 512
model = SIU3RModel(ModelCfg(image_size=(S, S)), precision="h3")
model.load_state_dict(synth.make_state_dict(populated=True))
model.cuda(local)
model.enable_cuda_graph()
img, K = synth.pair_inputs(1, 2, S)
img_pin, K_pin = img.pin_memory(), K.pin_memory()
img_d, K_d = img.cuda(), K.cuda()


def timed(fn, n):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn(n)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def value_loop(n):
    pend = None
    for i in range(n):
        h = model.forward_async(img_d, K_d, slot=i % 2)
        if pend is not None:
            model.forward_finish(pend, enable_query_class_logit_lift=True)
        pend = h
    model.forward_finish(pend, enable_query_class_logit_lift=True)


def make_e2e(fields, inputs):
    pipe = PairPipeline(model, lift=True, fields=fields)

    def run(n):
        for _ in range(n):
            pipe.submit(*inputs)
        pipe.flush()
    return run


out = {"rank": rank, "omp": torch.get_num_threads()}
value_loop(4)
out["value_ms"] = timed(value_loop, 20)
for name, fields, inputs in (("e2e_full", GAUSSIAN_FIELDS, (img_pin, K_pin)), ("e2e_no_d2h", (), (img_pin, K_pin)), ("e2e_dev_inputs", GAUSSIAN_FIELDS, (img_d, K_d)),
                             ("e2e_means_only", ("means",), (img_pin, K_pin))):
    run = make_e2e(fields, inputs)
    run(4)
    out[name + "_ms"] = timed(run, 20)
big = torch.empty(203_000_000 // 4, device="cuda")
bh = torch.empty(big.shape, pin_memory=True)
bh.copy_(big); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    bh.copy_(big, non_blocking=True)
torch.cuda.synchronize()
out["d2h_GBs"] = 5 * 0.203 / (time.perf_counter() - t0)
print(out, flush=True)
