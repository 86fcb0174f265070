"""SASS opcode histogram per kernel of libfitgnn_b200.so (cuobjdump -sass): which kernels use the tensor cores (UTCHMMA =
tcgen05.mma, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld), TMA (UTMALDG / UTMASTG, UBLKCP = cp.async.bulk), cp.async (LDGSTS),
clusters (UCGABAR), warp collectives, MUFU.  Usage: python scripts/sass_hist.py [lib.so] > profiles/rN_sass_hist.md"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "fitgnn_b200/libfitgnn_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", line)
    if m and kern:
        op = m.group(1)
        mod = m.group(2) or ""
        if op in ("UTCHMMA", "UTMALDG", "UTMASTG", "UTCBAR", "LDTM", "UBLKCP") and mod:
            op += "." + mod.strip(".").split(".")[0]
        hist[kern][op] += 1
demangled = subprocess.run(["c++filt"], input="\n".join(hist), capture_output=True, text=True).stdout.splitlines()
MARK = ("UTCHMMA", "UTMALDG", "UTMASTG", "UTCBAR", "LDTM", "UBLKCP", "LDGSTS", "UCGABAR", "SHFL", "MATCH", "REDUX", "MUFU",
        "LDS", "STS", "ATOM", "RED", "HMMA", "FFMA", "DFMA", "DMUL", "DADD")
print(f"# SASS opcode histogram of {lib} (sm_100a, cuobjdump -sass; static instruction counts)\n")
print("| kernel | instr | marked opcodes |")
print("|---|---|---|")
for (k, h), name in zip(hist.items(), demangled):
    short = re.sub(r"\(.*", "", name).replace("fitgnn::", "")
    marks = ", ".join(f"{op} {n}" for op, n in sorted(h.items()) if op.startswith(MARK))
    print(f"| `{short[:90]}` | {sum(h.values())} | {marks} |")
