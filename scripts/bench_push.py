"""Bandwidth of fitgnn_peer_push (bulk-copy push kernel) vs cudaMemcpyPeerAsync between two GPUs of one box, as a function of
the number of CTAs: python scripts/bench_push.py   (needs >= 2 visible GPUs; single process)"""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fitgnn_b200 as fg
from fitgnn_b200._lib import check, lib
assert torch.cuda.device_count() >= 2
n_peers = min(torch.cuda.device_count() - 1, 7)
nbytes = 57_600_000 // 16 * 16 if n_peers > 1 else 235_000_000 // 16 * 16
src = torch.rand(nbytes // 4, device="cuda:0")
dsts = [torch.zeros(nbytes // 4, device=f"cuda:{p + 1}") for p in range(n_peers)]
for d in dsts:  # torch enables peer access on the first cross-device copy
    d.copy_(src); src.copy_(d)
torch.cuda.set_device(0)
torch.cuda.synchronize()
def t(fn, reps=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / reps
arr = (C.c_void_p * n_peers)(*[C.c_void_p(d.data_ptr()) for d in dsts])
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for ctas in (1, 2, 4, 8, 16, 32):
    ms = t(lambda: check(lib().fitgnn_peer_push(C.c_void_p(src.data_ptr()), arr, n_peers, nbytes, ctas, st)))
    print(f"push kernel  {ctas:3d} CTAs -> {n_peers} peer(s): {ms:7.3f} ms  egress {n_peers * nbytes / ms / 1e6:7.1f} GB/s", flush=True)
for d in dsts:
    assert torch.equal(d.cpu(), src.cpu())
streams = [torch.cuda.Stream(device=0) for _ in range(n_peers)]
def ce():
    cur = torch.cuda.current_stream()
    for s_, d in zip(streams, dsts):
        s_.wait_stream(cur)
        with torch.cuda.stream(s_):
            d.copy_(src, non_blocking=True)
    for s_ in streams:
        cur.wait_stream(s_)
ms = t(ce)
print(f"copy engines, one stream per peer -> {n_peers} peer(s): {ms:7.3f} ms  egress {n_peers * nbytes / ms / 1e6:7.1f} GB/s")
