"""Static SASS opcode histogram per kernel: python scripts/sass_hist.py <obj-or-so> <substring of the mangled name>"""
import collections, re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
cur, hist = None, collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        cur = m.group(1); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
    if m and cur:
        toks = m.group(1).split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        hist[cur][op] += 1
for name, h in hist.items():
    if sys.argv[2] in name:
        print(name[:100], sum(h.values()))
        print("  " + ", ".join(f"{o} {n}" for o, n in h.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 22)))
