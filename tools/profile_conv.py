"""One tensor-core convolution (forward, optionally the data gradient) at a given shape inside a cudaProfilerStart/Stop window, with CUDA-event timings first.
    python tools/profile_conv.py Cin Cout H W [B] [--k 3] [--y16] [--relu] [--dgrad]
e.g. VGG19 conv1_2 of the perceptual branch: 64 64 256 512 32 --y16 --relu"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gan_danet_b200 import _lib, engine as E
ap = argparse.ArgumentParser()
ap.add_argument("cin", type=int); ap.add_argument("cout", type=int); ap.add_argument("H", type=int); ap.add_argument("W", type=int); ap.add_argument("B", type=int, nargs="?", default=32)
ap.add_argument("--k", type=int, default=3); ap.add_argument("--y16", action="store_true"); ap.add_argument("--relu", action="store_true"); ap.add_argument("--dgrad", action="store_true")
ap.add_argument("--no32", action="store_true", help="bf16 output only")
a = ap.parse_args()
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
B, H, W, Cin, Cout, k = a.B, a.H, a.W, a.cin, a.cout, a.k
x = torch.randn(B, H, W, Cin, generator=g).to(dev)
w = (0.05 * torch.randn(Cout, Cin, k, k, generator=g)).to(dev)
bias = torch.zeros(Cout, device=dev)
E.set_conv_precision("bf16")
xp = E.pack_act(x); del x
wf, wt = E.pack_weight(w, False), E.pack_weight(w, True)
y = None if a.no32 else torch.empty(B, H, W, Cout, device=dev)
y16 = torch.empty(B * H * W, Cout, dtype=torch.bfloat16, device=dev) if (a.y16 or a.no32) else None
act = _lib.ACT_RELU if a.relu else _lib.ACT_NONE
def fwd(): E.conv_tc_raw(xp, wf, y, (H, W), cin=Cin, kh=k, kw=k, pad=k // 2, bias=bias, act=act, y16=y16, out_shape=(B, H, W, Cout))
fns = [("fwd", fwd, 2.0 * B * H * W * Cin * Cout * k * k)]
if a.dgrad:
    dy = torch.randn(B, H, W, Cout, generator=g).to(dev); dyp = E.pack_act(dy); del dy
    gx = torch.empty(B, H, W, Cin, device=dev)
    fns.append(("dgrad", lambda: E.conv_tc_raw(dyp, wt, gx, (H, W), cin=Cout, kh=k, kw=k, pad=k // 2, transposed=True), 2.0 * B * H * W * Cin * Cout * k * k))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, fn, fl in fns:
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    print(f"{name}: {t:.4f} ms  {fl / t / 1e9:.0f} TFLOP/s  (C {Cin}->{Cout} k{k} {H}x{W} B{B} y16={bool(y16 is not None)} fp32={y is not None})", flush=True)
torch.cuda.profiler.start()
for _, fn, _ in fns: fn()
torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("done")
