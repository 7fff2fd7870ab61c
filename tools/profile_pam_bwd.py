"""One PAM forward + backward (C = 184, B = 32, N = 8192) with the backward inside a cudaProfilerStart/Stop window."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gan_danet_b200 import engine as E
from gan_danet_b200._lib import PREC_FP16
dev = torch.device("cuda:0")
B, H, W, C = 32, 64, 128, 184
d = C // 8
g = torch.Generator().manual_seed(0)
x = torch.randn(B, H, W, C, generator=g).to(dev)
q = (0.05 * torch.randn(B, H, W, d, generator=g)).to(dev)
k = (0.05 * torch.randn(B, H, W, d, generator=g)).to(dev)
v = torch.randn(B, H, W, C, generator=g).to(dev)
dy = torch.randn(B, H, W, C, generator=g).to(dev)
gamma = torch.full((1,), 0.5, device=dev)
for it in range(2):
    tape = E.Tape()
    xs = [E.Var(t) for t in (x, q, k, v)]
    y = E.op_pam_core(tape, *xs, E.Var(gamma), precision=PREC_FP16)
    y.g = dy
    torch.cuda.synchronize()
    if it == 1:
        torch.cuda.profiler.start()
    tape.backward()
    torch.cuda.synchronize()
    if it == 1:
        torch.cuda.profiler.stop()
print("done")
