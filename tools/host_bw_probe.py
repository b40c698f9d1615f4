"""Host-side ceiling of the end-to-end legs on a multi-GPU box: every rank copies 1 GiB pinned host <-> its GPU at the same
time (what N concurrent b200bgzf_*_host calls do, minus the kernels), and rank 0 prints the aggregate rates.
  python -m torch.distributed.run --nproc-per-node N tools/host_bw_probe.py        (N = 1: python tools/host_bw_probe.py)"""
import json, os, time
import torch
import torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("gloo")
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True); h.zero_()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    t = (time.perf_counter() - t0) / reps
    if world > 1:
        x = torch.tensor([t], dtype=torch.float64); dist.all_reduce(x, op=dist.ReduceOp.MAX); t = float(x.item())
    return t
res = {"gpus": world}
t = timed(lambda: d.copy_(h, non_blocking=True)); res["h2d_GBps_aggregate"] = round(world * n / t / 1e9, 1)
t = timed(lambda: h.copy_(d, non_blocking=True)); res["d2h_GBps_aggregate"] = round(world * n / t / 1e9, 1)
h2 = torch.empty(n // 4, dtype=torch.uint8, pin_memory=True); d2 = torch.empty(n // 4, dtype=torch.uint8, device="cuda"); s2 = torch.cuda.Stream()
def both():
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
    h.copy_(d, non_blocking=True)
t = timed(both); res["d2h_GBps_aggregate_with_quarter_h2d"] = round(world * n / t / 1e9, 1)
res["cpus"] = os.cpu_count()
if rank == 0: print(json.dumps(res))
if world > 1: dist.destroy_process_group()
