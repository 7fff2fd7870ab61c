"""BASELINE.json configs[3]: the 0.1-degree variant (generator grid 160x320 => PAM over N = 51200 positions, output 640x1280),
inference + backward on one B200.  Prints one JSON line:

  pam[C]      : fused PAM forward / backward launches (operand packing included) at N = 51200, algorithmic FLOPs
                2*B*N^2*(d+C) / 4*B*N^2*(d+C) over CUDA-event time
  inference   : FlexibleUpsamplingModule.forward in eval mode under no_grad (test.ipynb:143-167), ms per batch and samples/s
  fwd_bwd     : train-mode forward + backward of the generator (all parameter gradients), ms per batch

Inputs are resident in HBM; W >= 3 warm-up iterations; inputs + activations exceed L2 (trunk tensor alone 38 MB per sample,
full-resolution maps 210 MB) so no L2 flush is needed between iterations.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import gan_danet_b200 as P
from gan_danet_b200 import engine as E
from gan_danet_b200._lib import PREC_FP16

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--grid", default="160x320")
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--conv-precision", default="bf16")
args = ap.parse_args()
dev = torch.device("cuda:0")
H, W = (int(v) for v in args.grid.split("x"))
B, N = args.batch, H * W
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
peak_tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 0)) or 0)


def timeit(fn):
    for _ in range(args.warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.iters


out = {"workload": f"GAN-DANet generator, batch {B}, C_in 46, grid {H}x{W} (PAM over {N} positions), output {4 * H}x{4 * W}", "pam": {}}
for C in (160, 176, 184):
    d = C // 8
    g = torch.Generator().manual_seed(C)
    x, v, dy = (torch.randn(B, H, W, C, generator=g).to(dev) for _ in range(3))
    q, k = ((1.5 * torch.randn(B, H, W, d, generator=g)).to(dev) for _ in range(2))
    gamma = torch.full((1,), 0.5, device=dev)
    tf = timeit(lambda: E.op_pam_core(E.Tape(record=False), E.Var(x), E.Var(q), E.Var(k), E.Var(v), E.Var(gamma), precision=PREC_FP16))
    tape = E.Tape()
    xs = [E.Var(t) for t in (x, q, k, v)]
    gv = E.Var(gamma)
    y = E.op_pam_core(tape, *xs, gv, precision=PREC_FP16)
    y.g = dy

    def bwd():
        for t in xs + [gv]:
            t.g = None
        tape.ops[-1]()

    tb = timeit(bwd)
    ff, fb = 2.0 * B * N * N * (d + C), 4.0 * B * N * N * (d + C)
    out["pam"][str(C)] = {"fwd_ms": tf, "fwd_tflops": ff / tf / 1e9, "bwd_ms": tb, "bwd_tflops": fb / tb / 1e9,
                          "fwd_frac_of_peak": ff / tf / 1e9 / peak_tf if peak_tf else None}
    del x, v, dy, q, k, tape, xs, y

E.set_conv_precision(args.conv_precision)
torch.manual_seed(0)
G = P.FlexibleUpsamplingModule(46)
G.apply(P.weights_init_normal)
G = G.to(dev)
xin = torch.randn(B, 46, H, W, device=dev)
G.eval()


def infer():
    with torch.no_grad():
        return G(xin)


ti = timeit(infer)
out["inference"] = {"ms": ti, "samples_per_s": B / ti * 1e3}
G.train()
r = torch.randn(B, 1, 4 * H, 4 * W, device=dev)


def fwd_bwd():
    for p in G.parameters():
        p.grad = None
    G(xin).backward(r)


tt = timeit(fwd_bwd)
out["fwd_bwd"] = {"ms": tt, "samples_per_s": B / tt * 1e3}
out["peak_bf16_tflops"] = peak_tf
out["hbm_peak_allocated_gb"] = torch.cuda.max_memory_allocated() / 1e9
print(json.dumps(out))
