"""Per-kernel SASS mnemonic table of the built library (no GPU needed): proves which kernels use the Blackwell paths.
   python scripts/sass_table.py > profiles/r02_sass_table.txt
UTCHMMA = tcgen05.mma, .2CTA = cta_group::2, UTMALDG = TMA tensor load, UBLKCP = 1-D bulk copy, LDTM = tcgen05.ld,
UTCBAR = tcgen05.commit, SYNCS = mbarrier, FFMA2 / FADD2 / FMUL2 = packed fp32x2 arithmetic."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "nhans_b200", "libnhans_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UBLKCP", "LDTM", "UTCBAR", "SYNCS", "ELECT", "HMMA", "FFMA", "FFMA2", "FADD2", "FMUL2", "DFMA", "MUFU", "SHFL", "LDG", "STG", "LDS", "STS"]
tab = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"^void ", "", name).split("(")[0].replace("nhans::", "")
        cur = tab.setdefault(name, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]+)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["total"] += 1
        base = op.split(".")[0]
        if base in KEYS:
            cur[base] += 1
        if op.startswith("UTCHMMA") and ".2CTA" in op:
            cur["UTCHMMA.2CTA"] += 1
print("SASS mnemonic counts per kernel: cuobjdump -sass %s (sm_100a)" % os.path.relpath(so, ROOT))
print("%-62s %6s " % ("kernel", "instr") + " ".join("%7s" % k.replace("UTCHMMA.2CTA", ".2CTA") for k in KEYS))
agg = collections.OrderedDict()
for name, c in tab.items():
    base = re.sub(r"<.*", "", name)
    a = agg.setdefault(base, [0, collections.Counter()])
    a[0] += 1
    a[1].update(c)
for base, (n, c) in agg.items():
    label = "%s (%d instantiation%s)" % (base, n, "" if n == 1 else "s")
    print("%-62s %6d " % (label, c["total"]) + " ".join("%7d" % c[k] for k in KEYS))
