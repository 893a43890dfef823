"""torchrun --nproc-per-node N tools/nccl_probe.py : which transport NCCL picked and what the table exchange costs."""
import os, time, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"]); w = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 1 << 30
a = torch.zeros(n, dtype=torch.uint8, device="cuda"); b = torch.empty_like(a)
for name, fn in (("all_to_all_single 1GiB", lambda: dist.all_to_all_single(b, a)),
                 ("all_gather_into_tensor 1GiB", lambda: dist.all_gather_into_tensor(b, a[: n // w].clone())),
                 ("all_reduce max 1MiB", lambda: dist.all_reduce(a[: 1 << 20], op=dist.ReduceOp.MAX)),
                 ("all_reduce sum u8 1GiB", lambda: dist.all_reduce(a, op=dist.ReduceOp.SUM))):
    for _ in range(2): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    if rank == 0: print(f"{name}: {e0.elapsed_time(e1) / 5:.3f} ms", flush=True)
if rank == 0:
    print("can_device_access_peer(0,1):", torch.cuda.can_device_access_peer(0, 1), flush=True)
dist.destroy_process_group()
