"""Narrow-output weight gradient (conv_tc_wgrad_col_kernel) against the general tensor-core weight-gradient kernel and float64, with timings.
    python tools/check_wgrad_col.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from gan_danet_b200 import _lib, engine as E
dev = torch.device("cuda:0")
lib = _lib.lib_for_device(0)
E.set_conv_precision("bf16")
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
def run(B, H, W, Cin, Cout, timing=False):
    g = torch.Generator().manual_seed(Cin * 7 + H)
    x = torch.randn(B, H, W, Cin, generator=g).to(dev)
    dy = torch.randn(B, H, W, Cout, generator=g).to(dev)
    xp, dyp = E.pack_act(x), E.pack_act(dy)
    outs = {}
    for name, col, row in (("general", 0, 0), ("col", 1, 0), ("col_row", 1, 1)):
        lib.gdn_conv_tc_set_wgrad_col(col)
        lib.gdn_conv_tc_set_wgrad_col_row(row)
        gw = torch.zeros(Cout, Cin, 3, 3, device=dev)
        fn = lambda: E.wgrad_tc_raw(dyp, xp, gw, B=B, in_hw=(H, W), out_hw=(H, W), cin=Cin, cout=Cout, kh=3, kw=3, pad=1)
        fn(); torch.cuda.synchronize()
        t = None
        if timing:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): fn()
            e1.record(); torch.cuda.synchronize(); t = e0.elapsed_time(e1) / 10
        outs[name] = (gw.clone(), t)
    lib.gdn_conv_tc_set_wgrad_col(1); lib.gdn_conv_tc_set_wgrad_col_row(1)
    xr, dr = x.to(torch.bfloat16).double().permute(0, 3, 1, 2), dy.to(torch.bfloat16).double().permute(0, 3, 1, 2)
    w = torch.zeros(Cout, Cin, 3, 3, dtype=torch.float64, device=dev, requires_grad=True)
    (F.conv2d(xr, w, padding=1) * dr).sum().backward()
    ref = w.grad
    print(f"B{B} {H}x{W} C{Cin}->{Cout}: " + "  ".join(f"{k}: vs f64 {rel(v[0], ref):.2e}" + (f" {v[1]*1e3:.0f} us" if v[1] else "") for k, v in outs.items()), flush=True)
for case in ((1, 8, 16, 64, 24), (2, 16, 128, 136, 24), (1, 45, 22, 88, 24), (1, 5, 200, 88, 24), (2, 64, 128, 160, 24), (1, 32, 64, 80, 20), (1, 16, 128, 112, 8)):
    run(*case)
for Cin in (64, 88, 136, 160):
    run(32, 64, 128, Cin, 24, timing=True)


def run_dgrad(B, H, W, Cin, Cout, timing=False):
    """data gradient of the Cin -> Cout convolution (Cout narrow): COL mode against the per-tap kernel and float64"""
    g = torch.Generator().manual_seed(Cin * 3 + W)
    dy = torch.randn(B, H, W, Cout, generator=g).to(dev)
    w = (0.1 * torch.randn(Cout, Cin, 3, 3, generator=g)).to(dev)
    res0 = torch.randn(B, H, W, Cin, generator=g).to(dev)
    dyp, wt = E.pack_act(dy), E.pack_weight(w, True)
    outs = {}
    for name, col in (("per-tap", 0), ("col", 1)):
        lib.gdn_conv_tc_set_col(col)
        gx = res0.clone()
        fn = lambda: E.conv_tc_raw(dyp, wt, gx, (H, W), cin=Cout, kh=3, kw=3, pad=1, transposed=True, res=gx)
        fn(); torch.cuda.synchronize()
        out = gx.clone(); t = None
        if timing:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): fn()
            e1.record(); torch.cuda.synchronize(); t = e0.elapsed_time(e1) / 10
        outs[name] = (out, t)
    lib.gdn_conv_tc_set_col(1)
    ref = F.conv_transpose2d(dy.to(torch.bfloat16).double().permute(0, 3, 1, 2), w.to(torch.bfloat16).double(), padding=1).permute(0, 2, 3, 1) + res0.double()
    print(f"dgrad B{B} {H}x{W} C{Cout}->{Cin}: " + "  ".join(f"{k}: vs f64 {rel(v[0], ref):.2e}" + (f" {v[1]*1e3:.0f} us" if v[1] else "") for k, v in outs.items()), flush=True)
for case in ((1, 16, 128, 136, 24), (2, 3, 256, 112, 24), (1, 8, 128, 64, 20), (1, 5, 128, 160, 8), (1, 7, 128, 248, 24)):
    run_dgrad(*case)
for Cin in (64, 88, 136, 160):
    run_dgrad(32, 64, 128, Cin, 24, timing=True)
