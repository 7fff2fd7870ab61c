"""Micro-benchmark of the fused PAM kernels at BASELINE config 2 shapes (B = 32, N = 8192): forward and backward launches
(operand packing included, as in the step) against the algorithmic FLOP counts 2*B*N^2*(d+C) / 4*B*N^2*(d+C)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gan_danet_b200 import engine as E
from gan_danet_b200._lib import PREC_FP16

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--grid", default="64x128")
ap.add_argument("--channels", default="160,176,184")
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--no-bwd", action="store_true")
ap.add_argument("--qk-scale", type=float, default=1.5, help="std of q and k entries (1.5: logits of std ~11, rows near one-hot; 0.05: the flat rows of the network at init)")
args = ap.parse_args()
dev = torch.device("cuda:0")
H, W = (int(v) for v in args.grid.split("x"))
B, N = args.batch, H * W


def timeit(fn, iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for C in (int(c) for c in args.channels.split(",")):
    d = C // 8
    g = torch.Generator().manual_seed(C)
    x = torch.randn(B, H, W, C, generator=g).to(dev)
    q = (args.qk_scale * torch.randn(B, H, W, d, generator=g)).to(dev)
    k = (args.qk_scale * torch.randn(B, H, W, d, generator=g)).to(dev)
    v = torch.randn(B, H, W, C, generator=g).to(dev)
    dy = torch.randn(B, H, W, C, generator=g).to(dev)
    gamma = torch.full((1,), 0.5, device=dev)

    def fwd():
        t = E.Tape(record=False)
        return E.op_pam_core(t, E.Var(x), E.Var(q), E.Var(k), E.Var(v), E.Var(gamma), precision=PREC_FP16)

    tf = timeit(fwd, args.iters)
    line = f"C={C} d={d} B={B} N={N}: fwd {tf:7.3f} ms {2.0 * B * N * N * (d + C) / tf / 1e9:7.1f} TFLOP/s"
    if not args.no_bwd:
        tape = E.Tape()
        xs = [E.Var(t) for t in (x, q, k, v)]
        gv = E.Var(gamma)
        y = E.op_pam_core(tape, *xs, gv, precision=PREC_FP16)
        y.g = dy

        def bwd():
            for t in xs + [gv]:
                t.g = None
            tape.ops[-1]()

        tb = timeit(bwd, args.iters)
        line += f"   bwd {tb:7.3f} ms {4.0 * B * N * N * (d + C) / tb / 1e9:7.1f} TFLOP/s"
    print(line, flush=True)
