"""Micro-benchmark of the thin-convolution kernels at BASELINE config 2 shapes (B = 32, 256x512 fields, C = 64)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gan_danet_b200 import engine as E

dev = torch.device("cuda:0")


def timeit(fn, iters=5):
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


for name, (B, Cin, Cout, H, W, stride) in {"vgg1_1 1->64": (32, 1, 64, 256, 512, 1), "dconv1 1->64 s2": (64, 1, 64, 256, 512, 2), "final 64->1": (32, 64, 1, 256, 512, 1)}.items():
    Ho, Wo = (H + 2 - 3) // stride + 1, (W + 2 - 3) // stride + 1
    x = torch.randn(B, H, W, Cin, device=dev)
    w = torch.randn(Cout, Cin, 3, 3, device=dev) * 0.1
    y = torch.empty(B, Ho, Wo, Cout, device=dev)
    dy = torch.randn(B, Ho, Wo, Cout, device=dev)
    gx = torch.empty_like(x)
    gw = torch.empty_like(w)
    wide = max(x.numel(), y.numel()) * 4
    for thin in (True, False):
        E.thin_conv_enabled = thin
        E.set_conv_precision("bf16")
        ctx = E.conv_forward(x, w, y, stride=stride, pad=1)
        tf = timeit(lambda: E.conv_forward(x, w, y, stride=stride, pad=1))
        tw = timeit(lambda: E.conv_backward(ctx, dy, x, w, stride=stride, pad=1, gw=gw))
        td = timeit(lambda: E.conv_backward(ctx, dy, x, w, stride=stride, pad=1, gx=gx))
        print(f"{name:18s} thin={thin!s:5s} fwd {tf:6.3f} ms ({wide / tf / 1e6:5.0f} GB/s)  wgrad {tw:6.3f} ms  dgrad {td:6.3f} ms", flush=True)
