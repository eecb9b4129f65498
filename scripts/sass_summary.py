"""cuobjdump -sass vlm-fo1_b200/libfo1.so | python scripts/sass_summary.py > profiles/r02_sass_summary.txt
Per kernel: counts of the opcodes that prove which hardware path it uses."""
import collections, re, subprocess, sys
cols = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG.2D", "UTMALDG.3D", "UTMALDG.4D", "UTMALDG.2CTA", "UTCBAR.2CTA.MULTICAST", "UBLKCP", "LDTM", "STTM", "HMMA", "LDGSTS", "MUFU.EX2"]
cur = None
counts = collections.OrderedDict()
pat = re.compile(r"\b(UTCHMMA[.A-Z0-9]*|UTMALDG[.A-Z0-9]*|UTCBAR[.A-Z0-9]*|UBLKCP[.A-Z]*|LDTM[.A-Za-z0-9]*|STTM[.A-Za-z0-9]*|HMMA[.A-Z0-9]*|LDGSTS[.A-Z0-9]*|MUFU\.EX2)")
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); counts[cur] = collections.Counter(); continue
    if cur is None:
        continue
    for t in pat.findall(line):
        if t.startswith("UTCHMMA"): k = "UTCHMMA.2CTA" if "2CTA" in t else "UTCHMMA"
        elif t.startswith("UTMALDG"): k = "UTMALDG.2CTA" if "2CTA" in t else t[:10]
        elif t.startswith("UTCBAR"): k = "UTCBAR.2CTA.MULTICAST" if "2CTA" in t else None
        elif t.startswith("UBLKCP"): k = "UBLKCP"
        elif t.startswith("LDTM"): k = "LDTM"
        elif t.startswith("STTM"): k = "STTM"
        elif t.startswith("HMMA"): k = "HMMA"
        elif t.startswith("LDGSTS"): k = "LDGSTS"
        else: k = t
        if k: counts[cur][k] += 1
print("# SASS mnemonics per kernel of vlm-fo1_b200/libfo1.so (cuobjdump -sass, sm_100a, final round-2 build)")
print("# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2, the CTA-pair GEMM), UTMALDG = TMA tensor load (.2CTA = credits the pair leader's mbarrier),")
print("# UTCBAR.2CTA.MULTICAST = tcgen05.commit multicast to both CTAs, UBLKCP = cp.async.bulk, LDTM / STTM = tcgen05.ld / st, HMMA = mma.sync,")
print("# LDGSTS = cp.async.  Kernels without any of these (norms, rope, conv helpers, pre-processing ...) are omitted.")
print("kernel | " + " | ".join(cols))
for f, c in counts.items():
    if not c:
        continue
    name = subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void ", "")
    print(name + " | " + " | ".join(str(c.get(k, 0)) for k in cols))
