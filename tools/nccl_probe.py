"""All-reduce latency probe: torchrun --nproc-per-node N tools/nccl_probe.py"""
import os, time, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
for n, dt in ((131072, torch.float64), (1024, torch.int64), (12_500_000, torch.int32)):
    t = torch.ones(n, dtype=dt, device="cuda")
    for _ in range(3): dist.all_reduce(t)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): dist.all_reduce(t)
    torch.cuda.synchronize()
    if rank == 0: print(n, dt, "%.3f ms per all_reduce" % ((time.perf_counter() - t0) * 100))
dist.destroy_process_group()
