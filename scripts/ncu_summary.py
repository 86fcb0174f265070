"""Summarise an .ncu-rep (raw page) into a small table: python scripts/ncu_summary.py <rep> [out.md]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__shared_mem_per_block_dynamic"]
idx = [(w, hdr.index(w)) for w in want if w in hdr]
out = ["| " + " | ".join(w for w, _ in idx) + " |", "|" + "---|" * len(idx)]
out.append("| " + " | ".join(units[i] for _, i in idx) + " |")
for r in data:
    out.append("| " + " | ".join(r[i][:60] for _, i in idx) + " |")
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(f"# ncu summary of {rep}\n\n" + txt + "\n")
