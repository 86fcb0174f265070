import json, sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l)
        if "ms_per_step" in d:
            print("headline ms %.3f  nodes/s %.4g  e2e %s  parity %s" % (d["ms_per_step"], d["value"], d.get("e2e", {}).get("ms_per_step"), d.get("parity", {}).get("max_rel_err")))
            for k, v in d["kernels"].items():
                print(f"   {k:14s} {v['ms']:8.3f} ms {v['GBps']:8.1f} GB/s {v['TFLOPs']:6.1f} TF share {v['share']:.1%}")
            print("   roofline_spmm", round(d["roofline_spmm"]["frac"], 3), d["roofline_spmm"]["kernel"], "clocks", d["clocks"])
        for m, b in d.get("modes", {}).items():
            if m == "per_query":
                print("== per_query", {k_: v_ for k_, v_ in b.items() if k_ not in ("workload", "timer")})
                continue
            if m == "train":
                print("== train ms %.2f" % b["ms_per_step"], "nodes/s %.3g" % b["value"], "loss", b["loss_first_last"], "gemm_tn", b["gemm_tn"])
                continue
            print("==", m, "ms %.3f" % b["ms_per_step"], "nodes/s %.3g" % b["value"], "launches", b["gpu_launches"], "hub", b.get("hub_rows"), b.get("schedule"))
            print("   pack", b.get("pack"))
            for k, v in b["kernels"].items():
                print(f"   {k:14s} {v['ms']:8.3f} ms {v['GBps']:8.1f} GB/s {v['TFLOPs']:6.1f} TF share {v['share']:.1%} algoGB {v['algo_GB']:.2f}")
            if b.get("roofline_spmm"):
                print("   roofline_spmm", b["roofline_spmm"]["kernel"], round(b["roofline_spmm"]["frac"], 3))
            print("   parity", b["parity"]["max_rel_err"], b["parity"]["rows"])
    elif "rror" in l or "Traceback" in l or "PARITY" in l:
        print(l[:400])
