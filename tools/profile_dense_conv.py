"""Dense-block growth convolution (C_i -> 24, 3x3) forward and its data gradient (24 -> C_i) at B 32, 64x128, inside a cudaProfilerStart/Stop window.
    python tools/profile_dense_conv.py [Cin]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gan_danet_b200 import _lib, engine as E
dev = torch.device("cuda:0")
Cin = int(sys.argv[1]) if len(sys.argv) > 1 else 136
B, H, W, Cout = 32, 64, 128, 24
g = torch.Generator().manual_seed(0)
x = torch.randn(B, H, W, Cin, generator=g).to(dev)
dy = torch.randn(B, H, W, Cout, generator=g).to(dev)
w = (0.05 * torch.randn(Cout, Cin, 3, 3, generator=g)).to(dev)
E.set_conv_precision("bf16")
xp, dyp = E.pack_act(x), E.pack_act(dy)
wf, wt = E.pack_weight(w, False), E.pack_weight(w, True)
y = torch.empty(B, H, W, Cout, device=dev); gx = torch.empty(B, H, W, Cin, device=dev)
def fwd(): E.conv_tc_raw(xp, wf, y, (H, W), cin=Cin, kh=3, kw=3, pad=1)
def dgrad(): E.conv_tc_raw(dyp, wt, gx, (H, W), cin=Cout, kh=3, kw=3, pad=1, transposed=True)
for name, fn in (("fwd", fwd), ("dgrad", dgrad)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 10:.4f} ms (Cin {Cin})")
torch.cuda.profiler.start(); fwd(); dgrad(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("done")
