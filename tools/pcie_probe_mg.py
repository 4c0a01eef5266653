#!/usr/bin/env python3
"""Host <-> device ceiling of the box for the e2e leg when k = 1, 2, 4, 8 GPUs copy AT THE SAME TIME (one process per GPU under
torchrun, pinned 32 MB buffers, both directions on two streams): per-GPU and aggregate GB/s.  Ranks >= k idle at the barriers.
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_probe_mg.py"""
import os, time, json
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
N = 32 * 1024 * 1024
h_in = [torch.empty(N, dtype=torch.uint8).pin_memory() for _ in range(4)]
h_out = [torch.empty(N, dtype=torch.uint8).pin_memory() for _ in range(4)]
d_in = [torch.empty(N, dtype=torch.uint8, device="cuda") for _ in range(4)]
d_out = [torch.empty(N, dtype=torch.uint8, device="cuda") for _ in range(4)]
s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
def run(up, dn, reps=20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for r in range(reps):
        for i in range(4):
            if up:
                with torch.cuda.stream(s_up): d_in[i].copy_(h_in[i], non_blocking=True)
            if dn:
                with torch.cuda.stream(s_dn): h_out[i].copy_(d_out[i], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return reps * 4 * N / dt / 1e9
run(True, True, 2)
rows = []
k = 1
while k <= world:
    res = {}
    for name, (u, d) in (("h2d", (True, False)), ("d2h", (False, True)), ("both_each_dir", (True, True))):
        dist.barrier()
        v = run(u, d) if rank < k else 0.0
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        res[name] = {"aggregate_gbs": round(float(t.item()), 1), "per_gpu_gbs": round(float(t.item()) / k, 1)}
    if rank == 0:
        rows.append({"gpus_copying": k, **res})
        print(f"{k} GPU(s): H2D {res['h2d']['aggregate_gbs']:7.1f} GB/s  D2H {res['d2h']['aggregate_gbs']:7.1f}  both {res['both_each_dir']['aggregate_gbs']:7.1f} each way ({2 * res['both_each_dir']['aggregate_gbs']:.1f} total); per GPU {res['h2d']['per_gpu_gbs']} / {res['d2h']['per_gpu_gbs']} / {res['both_each_dir']['per_gpu_gbs']}", flush=True)
    k *= 2
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump({"what": "pinned host <-> device copies, k GPUs at the same time, one process per GPU", "cpus": os.cpu_count(), "rows": rows}, open("gpurun_out/pcie_probe_mg.json", "w"), indent=1)
dist.destroy_process_group()
