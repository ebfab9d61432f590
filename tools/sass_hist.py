"""Opcode histogram of the kernels of one generated module (cuobjdump -sass; no GPU needed): the evidence that
the module is a TMA + mbarrier kernel / holds no fused multiply-adds / uses 128-bit shared-memory loads.
python tools/sass_hist.py dawn_b200/lib/stencils/hori_diff_type2_stencil.so > profiles/r2_sass_hori_diff_type2.txt"""
import collections
import re
import subprocess
import sys

out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], stdout=subprocess.PIPE, text=True).stdout
kern, hist = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip().split("(")[0]
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
WATCH = ("UTMALDG", "UBLKCP", "SYNCS", "LDS", "STS", "LDG", "STG", "DFMA", "DMUL", "DADD", "FFMA", "FMUL", "FADD", "MUFU", "BAR", "SHFL",
         "MEMBAR", "ATOM", "LDL", "STL", "CCTL", "ERRBAR")
print("# " + " ".join(sys.argv))
for k, h in hist.items():
    print("kernel %s: %d instructions" % (k, sum(h.values())))
    groups = collections.Counter()
    for op, n in h.items():
        for w in WATCH:
            if op.startswith(w):
                groups[op] += n
    for op, n in sorted(groups.items()):
        print("  %-34s %6d" % (op, n))
    for w in ("DFMA", "FFMA", "UTMALDG", "SHFL", "LDL", "STL"):
        if not any(op.startswith(w) for op in h):
            print("  (no %s)" % w)
