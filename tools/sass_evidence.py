"""Static evidence from the built objects (no GPU needed): per kernel registers / shared memory / stack from `cuobjdump -res-usage` and counts
of the SASS mnemonics that prove the Blackwell paths (UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA,
HMMA = legacy mma.sync).  Writes a markdown table.

    python tools/sass_evidence.py > profiles/r01_sass_evidence.md
"""
from __future__ import annotations

import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "siu3r_b200", "csrc", "build")
PATTERNS = [("UTC*MMA", r"\bUTC\w*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"),
            ("SYNCS (mbarrier)", r"\bSYNCS"), ("HMMA", r"\bHMMA"), ("FFMA", r"\bFFMA"), ("MUFU", r"\bMUFU")]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(name: str) -> str:
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\(.*$", "", name)


def main():
    rows = []
    for obj in sorted(os.listdir(BUILD)):
        if not obj.endswith(".o"):
            continue
        path = os.path.join(BUILD, obj)
        res = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout
        usage = {}
        for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", res):
            usage[m.group(1)] = tuple(int(x) for x in m.groups()[1:])
        sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
        counts, cur = {}, None
        for line in sass.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                cur = m.group(1)
                counts[cur] = {k: 0 for k, _ in PATTERNS}
                continue
            if cur:
                for k, pat in PATTERNS:
                    if re.search(pat, line):
                        counts[cur][k] += 1
        names = demangle(list(usage))
        for fn, (reg, stack, shared, local) in usage.items():
            c = counts.get(fn, {k: 0 for k, _ in PATTERNS})
            rows.append((obj.replace(".o", ".cu"), short(names.get(fn, fn)), reg, stack, shared, c))
    print("# Static SASS / resource evidence of the built kernels (round 1)\n")
    print("`python tools/sass_evidence.py` over `siu3r_b200/csrc/build/*.o` (nvcc 12.9, `-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo`).")
    print("UTC*MMA = `tcgen05.mma`, LDTM / STTM = `tcgen05.ld` / `tcgen05.st`, UTMALDG / UTMASTG = TMA tensor loads / stores, SYNCS = mbarrier ops;")
    print("HMMA would be the legacy `mma.sync` path (none).  Static shared memory only (the tensor-core kernels use dynamic shared memory).\n")
    print("| file | kernel | regs | stack B | static smem B | " + " | ".join(k for k, _ in PATTERNS) + " |")
    print("|---|---|---:|---:|---:|" + "---:|" * len(PATTERNS))
    for f, k, reg, stack, shared, c in sorted(rows, key=lambda r: (r[0], r[1])):
        print(f"| {f} | `{k[:70]}` | {reg} | {stack} | {shared} | " + " | ".join(str(c[p]) for p, _ in PATTERNS) + " |")
    tc = [r for r in rows if r[5]["UTC*MMA"] > 0]
    legacy = sorted({r[1][:40] for r in rows if r[5]["HMMA"] > 0})
    print(f"\nLegacy `mma.sync` kernels: {legacy or 'none'} -- `flash_attn_d64_kernel` is the attention of the 3xTF32 parity mode (precision = 'fp32x3'); "
          "the TF32 mode that the bench measures runs `flash_tc_kernel` (tcgen05, P in TMEM) for every head-dim-64 attention.")
    print(f"\n{len(rows)} kernels; {len(tc)} issue tcgen05.mma; {sum(1 for r in rows if r[5]['UTMALDG'] > 0)} load through TMA; "
          f"{sum(1 for r in rows if r[5]['HMMA'] > 0)} use the legacy HMMA path; kernels with a stack frame: "
          f"{[r[1][:40] for r in rows if r[3] > 0] or 'none'}.")


if __name__ == "__main__":
    sys.exit(main())
