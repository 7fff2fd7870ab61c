"""2-rank probe: can ProcessGroupNCCL collectives (async_op=True + wait, as GradientAllReduce issues them) be captured in a CUDA graph here?
torchrun --nproc-per-node 2 tools/probe_nccl_graph.py [global|thread_local|relaxed]"""
import os
import sys
import time

import torch
import torch.distributed as dist

mode = sys.argv[1] if len(sys.argv) > 1 else "thread_local"
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank = dist.get_rank()
grp = dist.new_group() if (len(sys.argv) > 2 and sys.argv[2] == "own_group") else None     # a communicator used ONLY by the captured collectives
big = torch.full((64 << 20,), float(rank + 1), device=dev)
small = torch.full((1000,), float(rank + 1), device=dev)
for _ in range(2):                    # eager warm-up (communicator set-up happens here)
    dist.all_reduce(big, group=grp)
    dist.all_reduce(small, group=grp)
torch.cuda.synchronize()
big.fill_(rank + 1.0); small.fill_(rank + 1.0)
g = torch.cuda.CUDAGraph()
t0 = time.time()
with torch.cuda.graph(g, capture_error_mode=mode):
    y = big * 2.0
    h1 = dist.all_reduce(y, async_op=True, group=grp)
    z = small + 1.0
    h2 = dist.all_reduce(z, async_op=True, group=grp)
    w = z * 3.0 if False else small * 3.0      # independent work while the collectives are in flight
    h1.wait(); h2.wait()
    out = y.sum() + z.sum() + w.sum()
print(f"rank {rank}: captured in {time.time() - t0:.2f}s (mode {mode})", flush=True)
for i in range(3):
    g.replay()
torch.cuda.synchronize()
world = dist.get_world_size()
exp_y = 2.0 * sum(range(1, world + 1))
exp_z = sum(r + 2.0 for r in range(world))
print(f"rank {rank}: y[0]={float(y[0])} (expect {exp_y}) z[0]={float(z[0])} (expect {exp_z}) out={float(out):.1f}", flush=True)
print(f"rank {rank}: entering eager barrier on the default group", flush=True)
dist.barrier()
print(f"rank {rank}: barrier passed", flush=True)
t = torch.ones(4, device=dev)
dist.all_reduce(t)
torch.cuda.synchronize()
print(f"rank {rank}: eager all_reduce after the graph ok ({float(t[0])})", flush=True)
g.replay()
torch.cuda.synchronize()
print(f"rank {rank}: replay after eager ok", flush=True)
os._exit(0)
print(f"rank {rank}: done", flush=True)
