"""Kernel timeline of ONE graph replay of the forward (CUPTI through torch.profiler; structure only -- never a bench number).
usage: python tools/timeline.py [S] [precision] [out.json]
Prints per-stream busy time, the union busy time, and the idle gaps on the busiest stream; dumps (name, stream, start_us, dur_us) per kernel."""
import collections, json, os, re, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from siu3r_b200 import synth
from siu3r_b200.model import ModelCfg, SIU3RModel
from torch.profiler import profile, ProfilerActivity

S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
prec = sys.argv[2] if len(sys.argv) > 2 else "h3"
out = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/timeline.json"
model = SIU3RModel(ModelCfg(image_size=(S, S)), precision=prec)
model.load_state_dict(synth.make_state_dict(populated=True))
model.cuda()
model.enable_cuda_graph()
img, K = synth.pair_inputs(1, 2, S)
img, K = img.cuda(), K.cuda()
for _ in range(4):
    h = model.forward_async(img, K)
    torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    h = model.forward_async(img, K)
    torch.cuda.synchronize()
tmp = out + ".trace.json"
os.makedirs(os.path.dirname(out), exist_ok=True)
prof.export_chrome_trace(tmp)
tr = json.load(open(tmp))
os.remove(tmp)
rows = []
for e in tr["traceEvents"]:
    if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e:
        rows.append(dict(name=re.sub(r"\(.*", "", e["name"].replace("(anonymous namespace)::", "")).replace("void ", "")[:60], start=float(e["ts"]), dur=float(e["dur"]),
                         stream=e.get("args", {}).get("stream")))
rows.sort(key=lambda r: r["start"])
t0 = rows[0]["start"]
for r in rows:
    r["start"] -= t0
end = max(r["start"] + r["dur"] for r in rows)
print(f"{len(rows)} device activities, span {end / 1e3:.3f} ms")
# union busy
iv = sorted((r["start"], r["start"] + r["dur"]) for r in rows)
busy, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
gaps = []
for s, e in iv[1:]:
    if s > cur_e:
        busy += cur_e - cur_s
        gaps.append((s - cur_e, cur_e))
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
print(f"union busy {busy / 1e3:.3f} ms; idle (no kernel on any stream) {(end - busy) / 1e3:.3f} ms in {len(gaps)} gaps, mean {sum(g for g, _ in gaps) / max(len(gaps), 1):.2f} us")
agg = collections.OrderedDict()
for r in rows:
    a = agg.setdefault(r["name"], [0, 0.0])
    a[0] += 1; a[1] += r["dur"]
for n, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{n:62s} {c:5d} {d / 1e3:8.3f} ms  {d / c:7.1f} us")
per_stream = collections.Counter()
for r in rows:
    per_stream[r["stream"]] += r["dur"]
print("busy per stream (ms):", {k: round(v / 1e3, 3) for k, v in per_stream.items()})
print("sum of durations", sum(r["dur"] for r in rows) / 1e3, "ms")
# the first 80 activities with gaps, to see a ViT layer
prev_end = 0
for r in rows[:int(os.environ.get('TL_ROWS', '400'))]:
    print(f"{r['start']:9.1f} +{r['dur']:7.1f}  gap {r['start'] - prev_end:6.1f}  s{r['stream']}  {r['name']}")
    prev_end = max(prev_end, r["start"] + r["dur"])
os.makedirs(os.path.dirname(out), exist_ok=True)
json.dump(rows, open(out, "w"))
