#!/usr/bin/env python
"""Dev tool (torchrun): latency of an 8-byte NCCL all-reduce and of a small send/recv ring, for the
multi-GPU overhead budget of semb_cg."""
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
x = torch.ones(1, dtype=torch.float64, device="cuda")
for _ in range(20):
    dist.all_reduce(x)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 500
e0.record()
for _ in range(n):
    dist.all_reduce(x)
e1.record(); torch.cuda.synchronize()
t_ar = e0.elapsed_time(e1) / n * 1e3
buf = torch.ones(200000, dtype=torch.float64, device="cuda"); rb = torch.empty_like(buf)
nxt, prv = (rank + 1) % world, (rank - 1) % world
def ring():
    ops = [dist.P2POp(dist.isend, buf, nxt), dist.P2POp(dist.irecv, rb, prv)]
    for r in dist.batch_isend_irecv(ops): r.wait()
for _ in range(10): ring()
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
e0.record()
for _ in range(100): ring()
e1.record(); torch.cuda.synchronize()
t_sr = e0.elapsed_time(e1) / 100 * 1e3
if rank == 0:
    print(f"world {world}: all_reduce(8 B) {t_ar:.1f} us; ring send/recv of 1.6 MB {t_sr:.1f} us")
dist.destroy_process_group()
