"""Probe: does torch symmetric memory (CUDA VMM + NVLS multicast) work on this box?  torchrun --nproc-per-node 2"""
import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
try:
    t = symm_mem.empty(1 << 20, dtype=torch.float32, device=torch.device("cuda", local))
    h = symm_mem.rendezvous(t, dist.group.WORLD)
    print(rank, "rendezvous ok", "world", h.world_size, "multicast_ptr", hex(h.multicast_ptr), "ptrs", [hex(p) for p in h.buffer_ptrs],
          "has_multicast_support", getattr(symm_mem._SymmetricMemory, "has_multicast_support", lambda *a: "n/a")(torch._C._autograd.DeviceType.CUDA if False else __import__("torch").distributed.distributed_c10d.DeviceType.CUDA if False else None) if False else "", flush=True)
    t.fill_(rank + 1)
    h.barrier()
    other = h.get_buffer((rank + 1) % h.world_size, (8,), torch.float32)
    print(rank, "peer sees", other[:2].tolist(), flush=True)
    h.barrier()
except Exception as e:
    import traceback; traceback.print_exc()
    print(rank, "symm_mem failed:", repr(e), flush=True)
dist.destroy_process_group()
