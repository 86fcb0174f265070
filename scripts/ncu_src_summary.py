"""Summarise an `ncu --page source --csv` export: stall totals, samples grouped by execution count (= warp role / loop level),
and the hottest instructions.  python scripts/ncu_src_summary.py file.csv [top_n]"""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
for k, hi in enumerate(hdr_idx):
    h = rows[hi]
    ci = {n: i for i, n in enumerate(h)}
    end = hdr_idx[k + 1] - 1 if k + 1 < len(hdr_idx) else len(rows)
    data = [r for r in rows[hi + 1:end] if len(r) > ci["# Samples"]]
    print("kernel:", rows[hi - 1][1][:110] if hi > 0 else "?")
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    tot = sum(int(r[ci["# Samples"]]) for r in data)
    agg = {s: sum(int(r[ci[s]] or 0) for r in data) for s in stalls}
    print("samples", tot, "instructions", len(data), "warp-instr executed", sum(int(r[ci["Instructions Executed"]]) for r in data))
    print({k_[6:]: v for k_, v in sorted(agg.items(), key=lambda x: -x[1]) if v})
    by = defaultdict(lambda: [0, 0, defaultdict(int)])
    for r in data:
        b = by[int(r[ci["Instructions Executed"]])]
        b[0] += int(r[ci["# Samples"]]); b[1] += 1
        for s in stalls: b[2][s[6:]] += int(r[ci[s]] or 0)
    print("by execution count: exec, samples, #instr, top stalls")
    for ex, (sm, n, st) in sorted(by.items(), key=lambda x: -x[1][0])[:12]:
        print(" ", ex, sm, n, dict(sorted(st.items(), key=lambda x: -x[1])[:4]))
    print("hottest instructions: index, exec, sass, samples, top stalls")
    for r in sorted(data, key=lambda r: -int(r[ci["# Samples"]]))[:top_n]:
        st = {s[6:]: int(r[ci[s]]) for s in stalls if int(r[ci[s]] or 0) > 0}
        print(" ", data.index(r), r[ci["Instructions Executed"]], r[ci["Source"]].strip()[:60], r[ci["# Samples"]],
              dict(sorted(st.items(), key=lambda x: -x[1])[:3]))
