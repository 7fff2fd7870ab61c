"""Micro-benchmark of the tcgen05 convolution kernels at the layer shapes of BASELINE config 2 (B = 32).

    python tools/bench_conv.py [--only NAME] [--iters 10] [--precision bf16]

Prints achieved TFLOP/s (algorithmic 2*B*Ho*Wo*Cout*Cin*k*k) for forward, data gradient and weight gradient, timed with CUDA
events on the launching stream; operands are packed outside the timed region.
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gan_danet_b200 import engine as E

SHAPES = {  # name: (B, Cin, Cout, H, W, k, stride)
    "fuse2_368to184": (32, 368, 184, 64, 128, 3, 1),
    "vgg1_2_64to64": (32, 64, 64, 256, 512, 3, 1),
    "vgg2_2_128to128": (32, 128, 128, 128, 256, 3, 1),
    "vgg3_2_256to256": (32, 256, 256, 64, 128, 3, 1),
    "vgg4_1_256to512": (32, 256, 512, 32, 64, 3, 1),
    "dense_136to24": (32, 136, 24, 64, 128, 3, 1),
    "up4_64to64": (32, 64, 64, 128, 256, 3, 1),
    "dconv2_64to128_s2": (64, 64, 128, 128, 256, 3, 2),
    "qkv_184to184_1x1": (32, 184, 184, 64, 128, 1, 1),
    "dense_dgrad_24to136": (32, 24, 136, 64, 128, 3, 1),
    "dense_dgrad_24to88": (32, 24, 88, 64, 128, 3, 1),
    "fuse_dgrad_184to368": (32, 184, 368, 64, 128, 3, 1),
}
ap = argparse.ArgumentParser()
ap.add_argument("--only", default=None)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--fwd-only", action="store_true")
args = ap.parse_args()
E.set_conv_precision(args.precision)
dev = torch.device("cuda:0")


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for name, (B, Cin, Cout, H, W, k, s) in SHAPES.items():
    if args.only and args.only != name:
        continue
    pad = k // 2
    Ho, Wo = (H + 2 * pad - k) // s + 1, (W + 2 * pad - k) // s + 1
    x = torch.randn(B, H, W, Cin, device=dev)
    w = torch.randn(Cout, Cin, k, k, device=dev) * 0.05
    y = torch.empty(B, Ho, Wo, Cout, device=dev)
    dy = torch.randn(B, Ho, Wo, Cout, device=dev)
    gx = torch.empty(B, H, W, Cin, device=dev)
    gw = torch.empty_like(w)
    xp, wp, wtp, dyp = E.pack_act(x), E.pack_weight(w, False), E.pack_weight(w, True), E.pack_act(dy)
    flops = 2.0 * B * Ho * Wo * Cout * Cin * k * k
    t_f = timeit(lambda: E.conv_tc_raw(xp, wp, y, (H, W), cin=Cin, kh=k, kw=k, stride=s, pad=pad), args.iters)
    line = f"{name:22s} fwd {t_f:7.3f} ms {flops / t_f / 1e9:7.1f} TF/s"
    if not args.fwd_only:
        t_d = timeit(lambda: E.conv_tc_raw(dyp, wtp, gx, (Ho, Wo), cin=Cout, kh=k, kw=k, stride=s, pad=pad, transposed=True), args.iters)
        t_w = timeit(lambda: E.wgrad_tc_raw(dyp, xp, gw, B=B, in_hw=(H, W), out_hw=(Ho, Wo), cin=Cin, cout=Cout, kh=k, kw=k, stride=s, pad=pad), args.iters)
        t_p = timeit(lambda: E.pack_act(x), args.iters)
        line += f" | dgrad {t_d:7.3f} ms {flops / t_d / 1e9:7.1f} TF/s | wgrad {t_w:7.3f} ms {flops / t_w / 1e9:7.1f} TF/s | pack(x) {t_p:6.3f} ms {x.numel() * 6 / t_p / 1e6:6.0f} GB/s"
    print(line, flush=True)
