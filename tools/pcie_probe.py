#!/usr/bin/env python3
"""PCIe ceiling of the box for the e2e leg: pinned host <-> device copies of 4K-picture-sized buffers, one direction and
both directions at once (two streams), GB/s."""
import torch, time
dev = torch.device("cuda", 0)
N = 32 * 1024 * 1024  # 32 MB
h_in = [torch.empty(N, dtype=torch.uint8).pin_memory() for _ in range(4)]
h_out = [torch.empty(N, dtype=torch.uint8).pin_memory() for _ in range(4)]
d_in = [torch.empty(N, dtype=torch.uint8, device=dev) for _ in range(4)]
d_out = [torch.empty(N, dtype=torch.uint8, device=dev) for _ in range(4)]
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
print(f"H2D alone  {run(True, False):6.1f} GB/s")
print(f"D2H alone  {run(False, True):6.1f} GB/s")
b = run(True, True)
print(f"both       {b:6.1f} GB/s each direction ({2*b:.1f} total)")
