"""SASS opcode histogram per kernel of the in-tree library (cuobjdump -sass): evidence that the hot kernels are tcgen05 / TMEM /
TMA code (UTCHMMA = tcgen05.mma, UTMALDG = cp.async.bulk.tensor, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit) and
not recompiled mma.sync (HMMA) kernels.   python tools/sass_histogram.py > profiles/r02_sass_histogram.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "physdock_b200", "csrc", "libphysdock_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEY = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "ELECT", "HMMA", "MUFU",
       "FADD2", "FFMA2", "FMUL2", "F2FP", "LDS", "STS", "LDG", "STG", "SHFL", "BAR", "DFMA", "DADD", "DMUL"]
cur, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        hist[cur][m.group(1)] += 1
demangled = subprocess.run(["c++filt"], input="\n".join(hist), capture_output=True, text=True).stdout.splitlines()
print(f"# SASS opcode histogram of {os.path.relpath(lib, ROOT)} (sm_100a), one row per kernel; columns = static instruction counts")
print(f"# {'kernel':70s} {'total':>6s} " + " ".join(f"{k:>8s}" for k in KEY))
tot = collections.Counter()
for (name, h), dn in zip(hist.items(), demangled):
    short = re.sub(r"\(anonymous namespace\)::", "", dn)
    short = re.sub(r"^void ", "", short).split("(")[0].replace("pdk::", "")
    n = sum(h.values())
    if n == 0:
        continue
    tot.update(h)
    print(f"  {short[:70]:70s} {n:6d} " + " ".join(f"{sum(v for k2, v in h.items() if k2.startswith(k)):8d}" for k in KEY))
print(f"  {'ALL KERNELS':70s} {sum(tot.values()):6d} " + " ".join(f"{sum(v for k2, v in tot.items() if k2.startswith(k)):8d}" for k in KEY))
