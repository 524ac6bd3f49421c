"""SASS census of the shipped library and of the run-time specialised cubins: per kernel, the instructions that tell which
hardware path it uses (DMMA = fp64 tensor pipe, LDGSTS = cp.async operand ring, UTMALDG / SYNCS = TMA + mbarrier, ...).
Usage: python scratch/sass_summary.py > profiles/r02_sass_summary.txt"""
import glob, os, re, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["DMMA", "DFMA", "LDGSTS", "UTMALDG", "SYNCS", "MUFU", "BAR.SYNC", "WARPSYNC", "SHFL", "ATOM", "RED", "LDL", "STL"]

def census(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    per, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            per[cur]["total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    per[cur][k] += 1
    return per

def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]
    except Exception:
        return n

files = [os.path.join(ROOT, "waveome_b200/_lib/libwaveome_b200.so")] + sorted(glob.glob(os.path.join(ROOT, "waveome_b200/_lib/rtc_cache/*.cubin")))
print("SASS census (cuobjdump -sass, sm_100a).  DMMA = fp64 tensor-core MMA (tcgen05 has no f64 kind), LDGSTS = cp.async,")
print("UTMALDG = cp.async.bulk.tensor (TMA), SYNCS = mbarrier ops, LDL/STL = local-memory (spill / indexed array) traffic.\n")
for f in files:
    per = census(f)
    print("== %s (%d kernels)" % (os.path.relpath(f, ROOT), len(per)))
    print("%-44s %7s " % ("kernel", "instrs") + " ".join("%8s" % k for k in KEYS))
    for name, c in per.items():
        print("%-44s %7d " % (demangle(name)[:44], c["total"]) + " ".join("%8d" % c[k] for k in KEYS))
    print()
