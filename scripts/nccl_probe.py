import os, time, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
for mb in (1, 16, 58, 230):
    n = mb * 1024 * 1024 // 4
    buf = torch.zeros(world, n, device=dev)
    for _ in range(3): dist.all_gather_into_tensor(buf.view(-1), buf[rank])
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): dist.all_gather_into_tensor(buf.view(-1), buf[rank])
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    if rank == 0: print(f"all_gather {mb} MB/rank x {world}: {ms:.3f} ms  -> recv {mb*(world-1)/ms:.1f} GB/s per rank", flush=True)
# raw peer copy
if world >= 2:
    can = torch.cuda.can_device_access_peer(lr, (lr + 1) % world)
    if rank == 0: print("can_device_access_peer", can, flush=True)
dist.destroy_process_group()
