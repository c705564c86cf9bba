"""Summarise an `ncu --csv --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]` launch list per kernel.
usage: python tools/summarize_launches.py launches.csv [--md] [--json out.json]"""
import collections, csv, json, re, sys


def load(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    hdr = rows[0]
    ki, mi, vi, ui, ii = (hdr.index(n) for n in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
    per = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        d = per.setdefault(r[ii], {"name": re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("<unnamed>::", "")})
        v, u = float(r[vi].replace(",", "")), r[ui]
        if r[mi].startswith("gpu__time_duration"):
            d["ms"] = v / 1e6 if u.startswith("n") else v / 1e3 if u.startswith("u") else v * (1e3 if u in ("s", "second") else 1)
        elif r[mi].startswith("dram__bytes"):
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            d["dram"] = d.get("dram", 0.0) + v * scale
    return list(per.values())


def main():
    path = sys.argv[1]
    ls = load(path)
    agg = collections.OrderedDict()
    for d in ls:
        a = agg.setdefault(d["name"][:64], {"launches": 0, "ms": 0.0, "dram": 0.0})
        a["launches"] += 1
        a["ms"] += d.get("ms", 0.0)
        a["dram"] += d.get("dram", 0.0)
    tot = sum(a["ms"] for a in agg.values())
    has_dram = any(a["dram"] for a in agg.values())
    print(f"Total {tot:.2f} ms over {len(ls)} launches.\n")
    print("| kernel | launches | ms | share |" + (" DRAM MB/launch | DRAM GB/s |" if has_dram else ""))
    print("|---|---:|---:|---:|" + ("---:|---:|" if has_dram else ""))
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        extra = f" {a['dram'] / a['launches'] / 1e6:.2f} | {a['dram'] / max(a['ms'], 1e-9) / 1e6:.0f} |" if has_dram else ""
        print(f"| `{n}` | {a['launches']} | {a['ms']:.3f} | {100 * a['ms'] / tot:.1f}% |{extra}")
    if "--traffic" in sys.argv:   # per-family DRAM bytes per launch -> profiles/r01_ncu_traffic.json (read by bench.py)
        fams = {"gemm_tc": ("gemm_tc",), "flash_attn": ("flash_tc", "flash_attn"), "conv2d_tc": ()}
        out = {}
        for fam, keys in fams.items():
            sel = [a for n, a in agg.items() if any(k in n for k in keys)]
            if sel:
                nl = sum(a["launches"] for a in sel)
                out[fam] = {"dram_bytes_per_launch": sum(a["dram"] for a in sel) / nl, "launches": nl, "ms": sum(a["ms"] for a in sel),
                            "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum over tools/one_forward.py 512 tf32 "
                                      "(all gemm_tc* launches: linear GEMMs and implicit-GEMM convs share the kernels)"}
        json.dump(out, open(sys.argv[sys.argv.index("--traffic") + 1], "w"), indent=1)
    if "--json" in sys.argv:
        json.dump({"total_ms": tot, "kernels": agg}, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
