"""Per-kernel SASS mnemonic counts of the built objects (tcgen05 / TMA / TMEM / legacy mma.sync evidence) -> markdown table.
usage: python tools/sass_evidence.py > profiles/r02_sass_evidence.md"""
import collections, glob, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMAPF", "SYNCS", "ELECT", "HMMA", "MUFU.EX2", "F2F.F16.F32", "F2FP", "ACQBULK", "UCGABAR"]
print("# SASS evidence, round 2 (cuobjdump -sass of siu3r_b200/csrc/build/*.o, sm_100a)\n")
print("Columns = occurrences of the mnemonic in the kernel's SASS.  UTCHMMA = tcgen05.mma kind::f16 / kind::tf32, LDTM / STTM = tcgen05.ld / .st (TMEM),")
print("UTMALDG = cp.async.bulk.tensor (TMA), SYNCS = mbarrier ops, ELECT = elect.sync (single-lane issue inside warp-uniform loops), HMMA = legacy mma.sync,")
print("ACQBULK = griddepcontrol.wait (programmatic dependent launch).  Kernels without any of these are omitted.\n")
print("| object | kernel | " + " | ".join(KEYS) + " |")
print("|---|---|" + "---:|" * len(KEYS))
for obj in sorted(glob.glob(os.path.join(ROOT, "siu3r_b200/csrc/build/*.o"))):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur, counts = None, collections.OrderedDict()
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            for k in KEYS:
                if op.startswith(k):
                    counts[cur][k] += 1
    for fn, c in counts.items():
        if not c:
            continue
        dem = subprocess.run(["cu++filt", fn], capture_output=True, text=True).stdout.strip() or fn
        dem = re.sub(r"\(.*", "", dem).replace("(anonymous namespace)::", "").replace("void ", "")
        print(f"| {os.path.basename(obj)} | `{dem[:70]}` | " + " | ".join(str(c.get(k, 0)) for k in KEYS) + " |")
