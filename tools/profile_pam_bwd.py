"""One PAM forward + backward (C = 184, B = 32, N = 8192, logits at the reference's scale) inside a cudaProfilerStart/Stop window.
    python tools/profile_pam_bwd.py [fp16x3|fp16]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gan_danet_b200 import engine as E
dev = torch.device("cuda:0")
B, H, W, C = 32, 64, 128, 184
d = C // 8
g = torch.Generator().manual_seed(0)
x = torch.randn(B, H, W, C, generator=g).to(dev)
q = (1.44 * torch.randn(B, H, W, d, generator=g)).to(dev)
k = (1.44 * torch.randn(B, H, W, d, generator=g)).to(dev)
v = torch.randn(B, H, W, C, generator=g).to(dev)
dy = (1e-4 * torch.randn(B, H, W, C, generator=g)).to(dev)
prec = E.PAM_PRECISION_NAMES[sys.argv[1] if len(sys.argv) > 1 else "fp16x3"]
gamma = torch.full((1,), 0.5, device=dev)
for it in range(2):
    tape = E.Tape()
    xs = [E.Var(t) for t in (x, q, k, v)]
    torch.cuda.synchronize()
    if it == 1:
        torch.cuda.profiler.start()
    y = E.op_pam_core(tape, *xs, E.Var(gamma), precision=prec)
    y.g = dy
    tape.backward()
    torch.cuda.synchronize()
    if it == 1:
        torch.cuda.profiler.stop()
print("done")
