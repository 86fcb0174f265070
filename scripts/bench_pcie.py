"""PCIe copy micro-benchmark for the e2e leg: H2D of the feature rows (980 MB) on 1/2/4 streams, alone and with the
concurrent D2H of the logits (470 MB)."""
import torch
dev = torch.device("cuda:0")
NB_IN, NB_OUT = 979_611_600, 470_213_568
h_in = torch.empty(NB_IN, dtype=torch.uint8).pin_memory()
d_in = torch.empty(NB_IN, dtype=torch.uint8, device=dev)
h_out = torch.empty(NB_OUT, dtype=torch.uint8).pin_memory()
d_out = torch.empty(NB_OUT, dtype=torch.uint8, device=dev)


def run(n_streams, with_d2h, reps=5):
    streams = [torch.cuda.Stream() for _ in range(n_streams)]
    s_out = torch.cuda.Stream()
    chunk = (NB_IN + n_streams - 1) // n_streams
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for s in streams + [s_out]:
            s.wait_event(a)
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                d_in[i * chunk:(i + 1) * chunk].copy_(h_in[i * chunk:(i + 1) * chunk], non_blocking=True)
        if with_d2h:
            with torch.cuda.stream(s_out):
                h_out.copy_(d_out, non_blocking=True)
        for s in streams + [s_out]:
            torch.cuda.current_stream().wait_stream(s)
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print(f"H2D {NB_IN / 1e6:.0f} MB on {n_streams} stream(s){' + concurrent D2H 470 MB' if with_d2h else ''}: "
          f"{best:.2f} ms = {NB_IN / best / 1e6:.1f} GB/s H2D", flush=True)


for ns in (1, 2, 4):
    run(ns, False)
for ns in (1, 2, 4):
    run(ns, True)
