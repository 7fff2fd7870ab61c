"""BASELINE.json configs[4]: deep ensemble of GAN-DANet generators, one member per B200, full-domain monthly TWSA sweep.

    python tools/bench_ensemble.py [--grid 45x22] [--months 181] [--batch 32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/bench_ensemble.py

Every rank holds ONE member (seed 42 + rank, deep_ensemble.ipynb:312) and sweeps the same T monthly fields
(deep_ensemble.ipynb:386-424) in eval mode: input preparation, generator, inverse scaling, plateau-masked spatial means on the
device; then the statistic exchange (all_gather of the member time series + of the fields, mean / std over the members).
Rank 0 prints one JSON line: fields/s of the whole job (members x months / max-over-ranks device time), eager launches and
CUDA-graph replay, with inputs resident in HBM and end to end from pinned host memory.  Weak scaling: per-GPU work is fixed.
"""
import argparse
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import gan_danet_b200 as P
from gan_danet_b200 import _lib as L
from gan_danet_b200 import engine as E
from gan_danet_b200.ensemble import EnsembleTrainer
from gan_danet_b200.synthetic import fast_batch

ap = argparse.ArgumentParser()
ap.add_argument("--grid", default="45x22", help="generator-input grid h x w (authors' sweep: 45x22 -> 180x88; BASELINE grid: 64x128)")
ap.add_argument("--months", type=int, default=181)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--conv-precision", default="bf16")
args = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
h, w = (int(v) for v in args.grid.split("x"))
T, B = args.months, args.batch
E.set_conv_precision(args.conv_precision)

ens = EnsembleTrainer(world, {}, ensemble_dir=tempfile.mkdtemp())
ens.set_seed(ens.seeds[rank])
G = P.FlexibleUpsamplingModule(46, attention_type="danet")
G.apply(P.weights_init_normal)
with torch.no_grad():
    for n, p in G.named_parameters():
        if n.endswith("gamma"):
            p.fill_(0.05)
G = G.to(dev).eval()
host = []
for t0 in range(0, T, B):
    lr05, lr025, aux = fast_batch(7 + t0, min(B, T - t0), h, w)
    host.append(tuple(t.pin_memory() for t in (lr05, lr025, aux)))
resident = [tuple(t.to(dev) for t in b) for b in host]
keep = (torch.rand(4 * h, 4 * w, generator=torch.Generator().manual_seed(1)) > 0.3).to(dev)
scaler = (9.5, 1.25)


def sweep(batches, use_graph):
    preds, trues = ens.predict_ensemble([G], batches, scaler=scaler, use_graph=use_graph)
    mean_ts, std_ts, r2 = ens.compute_uncertainty(preds, trues, keep)            # gathers [M, T, 1] (reads back T floats: the product)
    pm, ps = ens.pixel_statistics(preds)                                         # gathers the fields over NVLink
    return pm, ps, r2


def timed(batches, use_graph):
    for _ in range(args.warmup):
        sweep(batches, use_graph)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    before = L.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        sweep(batches, use_graph)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.iters], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms), (L.launch_count - before) // args.iters


out = {"metric": "ensemble sweep fields/s", "unit": "fields/s", "n_gpus": world, "scaling": "weak", "data": "synthetic",
       "config": {"workload": f"deep ensemble, {world} member(s) one per GPU, {T} monthly fields of [46,{h},{w}] -> [1,{4 * h},{4 * w}], batch {B}, eval mode, "
                              f"inverse scaling + masked spatial means + member mean/std on device", "conv_precision": args.conv_precision}}
for name, batches, graph in (("eager", resident, False), ("graph", resident, True), ("e2e_graph", host, True)):
    ms, calls = timed(batches, graph)
    out[name] = {"ms_per_sweep": ms, "fields_per_s": world * T / ms * 1e3, "abi_calls_per_sweep": calls}
out["value"] = out["graph"]["fields_per_s"]
bytes_in = sum(t.numel() * 4 for b in host for t in b)
out["e2e"] = {"value": out["e2e_graph"]["fields_per_s"], "unit": "fields/s", "h2d_bytes_per_step": bytes_in, "d2h_bytes_per_step": 2 * T * 4}
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
