"""Weight gradient of a dense-block growth convolution (C_i -> 24, 3x3, B 32, 64x128) inside a cudaProfilerStart/Stop window, role-swapped and direct.
    python tools/profile_wgrad.py [Cin]"""
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
E.set_conv_precision("bf16")
lib = _lib.lib_for_device(0)
xp, dyp = E.pack_act(x), E.pack_act(dy)
gw = torch.empty(Cout, Cin, 3, 3, device=dev)
def run():
    E.wgrad_tc_raw(dyp, xp, gw, B=B, in_hw=(H, W), out_hw=(H, W), cin=Cin, cout=Cout, kh=3, kw=3, stride=1, pad=1)
for swap in (1, 0):
    lib.gdn_conv_tc_set_wgrad_swap(swap)
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    print(f"swap={swap}: {e0.elapsed_time(e1) / 10:.4f} ms per weight gradient (Cin {Cin})")
    torch.cuda.profiler.start(); run(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("done")
