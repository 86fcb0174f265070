"""Radix-sort micro-benchmark (primitives.cu): random keys, (hi<<32|lo) keys and sentinel-heavy keys."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fitgnn_b200 as fg

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 123_000_000
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)


def timed(name, keys, bits, reps=3):
    ts = []
    for _ in range(reps):
        k = keys.clone()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fg.ops.sort_u64(k, None, key_bits=bits)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    passes = (bits + 7) // 8
    print(f"{name:28s} n={n:.3g} bits={bits} passes={passes}  {min(ts):8.3f} ms  = {min(ts) / passes:6.3f} ms/pass  "
          f"({3 * 8 * n / (min(ts) / passes * 1e-3) / 1e9:7.1f} GB/s per pass)", flush=True)
    assert bool((k[1:] >= k[:-1]).all())


keys = torch.randint(0, 2 ** 42, (n,), generator=g, device=dev, dtype=torch.int64)
timed("uniform 42-bit", keys, 42)
timed("uniform 16 of 42 bits", keys & 0xFFFF, 16)
sent = torch.where(torch.rand(n, device=dev, generator=g) < 0.95, torch.full_like(keys, (2 ** 21 - 1) << 21), keys)
timed("95% one sentinel key", sent, 42)
t0 = time.perf_counter()
torch.sort(keys)
torch.cuda.synchronize()
t0 = time.perf_counter()
torch.sort(keys)
torch.cuda.synchronize()
print(f"torch.sort (CUB, int64 + index)  {(time.perf_counter() - t0) * 1e3:8.3f} ms")
