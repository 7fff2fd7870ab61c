"""Backward of one DenseNet growth convolution (C_i -> 24, 3x3; generator.py:34) at B 32, 64x128: the narrow-output weight-gradient kernel and the COL-mode
data gradient, timed with CUDA events and then run once inside a cudaProfilerStart/Stop window (ncu --profile-from-start off).
    python tools/profile_dense_bwd.py [Cin]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gan_danet_b200 import engine as E
dev = torch.device("cuda:0")
Cin = int(sys.argv[1]) if len(sys.argv) > 1 else 136
B, H, W, Cout = 32, 64, 128, 24
g = torch.Generator().manual_seed(0)
x = torch.randn(B, H, W, Cin, generator=g).to(dev)
dy = torch.randn(B, H, W, Cout, generator=g).to(dev)
w = (0.05 * torch.randn(Cout, Cin, 3, 3, generator=g)).to(dev)
E.set_conv_precision("bf16")
xp, dyp, wt = E.pack_act(x), E.pack_act(dy), E.pack_weight(w, True)
gw = torch.empty_like(w); gx = torch.zeros(B, H, W, Cin, device=dev)
def wgrad(): E.wgrad_tc_raw(dyp, xp, gw, B=B, in_hw=(H, W), out_hw=(H, W), cin=Cin, cout=Cout, kh=3, kw=3, pad=1)
def dgrad(): E.conv_tc_raw(dyp, wt, gx, (H, W), cin=Cout, kh=3, kw=3, pad=1, transposed=True, res=gx)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, fn in (("wgrad (conv_tc_wgrad_col_kernel + reduce)", wgrad), ("dgrad (conv_tc_fwd_kernel<2>, accumulating)", dgrad)):
    for _ in range(3): fn()
    ts = []
    for _ in range(10):
        flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(f"{name}: {sorted(ts)[5] * 1e3:.0f} us (C {Cin} <-> {Cout}, B {B}, {H}x{W}, L2 flushed between runs)", flush=True)
torch.cuda.profiler.start(); wgrad(); dgrad(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("done")
