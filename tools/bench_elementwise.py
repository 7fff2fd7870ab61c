"""Achieved HBM bandwidth of the elementwise / reduction / resampling kernels at BASELINE config 2 shapes.
Algorithmic bytes = every tensor read once + written once (fp32)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gan_danet_b200 import engine as E, _lib as L
from gan_danet_b200._lib import ACT_RELU

dev = torch.device("cuda:0")
lib = L.lib_for_device(0)
S = lambda: E._stream()


def timeit(fn, iters=10):
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


def report(name, ms, nbytes):
    print(f"{name:34s} {ms:7.3f} ms  {nbytes / ms / 1e6:7.0f} GB/s", flush=True)


for (B, H, W, Cc, pitch) in [(32, 64, 128, 184, 184), (32, 64, 128, 136, 160), (32, 256, 512, 64, 64)]:
    M = B * H * W
    print(f"--- [{M} x {Cc}] pitch {pitch}")
    buf = torch.randn(M, pitch, device=dev)
    x = buf[:, :Cc]
    y = torch.empty(M, Cc, device=dev)
    dy = torch.randn(M, Cc, device=dev)
    n = M * Cc * 4
    report("colstats", timeit(lambda: E.colstats(x)), n)
    coef = torch.rand(4, Cc, device=dev) + 0.5
    report("affine_act", timeit(lambda: E.affine_act(x, y, coef[2], coef[3], ACT_RELU)), 2 * n)
    report("act_bwd", timeit(lambda: E.act_bwd(dy, y, y, ACT_RELU, 0.0)), 3 * n)
    report("axpy (accumulate)", timeit(lambda: E.axpy(dy, y, 1.0, True)), 3 * n)
    sums = torch.empty(2 * Cc, dtype=torch.float64, device=dev)
    ws = E.workspace("stat", lib.gdn_colstats_ws_bytes(M, Cc), dev)
    report("bn_bwd_reduce", timeit(lambda: L.check(lib.gdn_bn_bwd_reduce(dy.data_ptr(), Cc, 0, x.data_ptr(), pitch, 0, M, Cc, coef[0].data_ptr(), coef[1].data_ptr(),
           coef[2].data_ptr(), coef[3].data_ptr(), ACT_RELU, 0.0, sums.data_ptr(), ws.data_ptr(), S()))), 2 * n)
    report("bn_bwd_apply", timeit(lambda: L.check(lib.gdn_bn_bwd_apply(dy.data_ptr(), Cc, 0, x.data_ptr(), pitch, 0, y.data_ptr(), Cc, 0, 0, M, Cc, coef[0].data_ptr(), coef[1].data_ptr(),
           coef[2].data_ptr(), coef[2].data_ptr(), coef[3].data_ptr(), ACT_RELU, 0.0, sums.data_ptr(), None, None, S()))), 3 * n)
    report("pack_act (bf16)", timeit(lambda: E.pack_act(x)), n * 3 // 2)
    del buf, x, y, dy

B, H, W, Cc = 32, 128, 256, 64
x = torch.randn(B, H, W, Cc, device=dev)
yb = torch.empty(B, 2 * H, 2 * W, Cc, device=dev)
n_in, n_out = x.numel() * 4, yb.numel() * 4
report("bicubic_up2_fwd 128x256->256x512", timeit(lambda: L.check(lib.gdn_bicubic_up2_fwd(x.data_ptr(), yb.data_ptr(), B, H, W, Cc, S()))), n_in + n_out)
report("bicubic_up2_bwd", timeit(lambda: L.check(lib.gdn_bicubic_up2_bwd(yb.data_ptr(), x.data_ptr(), B, H, W, Cc, S()))), n_in + n_out)
xs = torch.randn(B, 64, 128, Cc, device=dev)
report("bilinear_fwd 64x128->256x512 (+=)", timeit(lambda: L.check(lib.gdn_bilinear_fwd(xs.data_ptr(), yb.data_ptr(), B, 64, 128, 256, 512, Cc, 1, S()))), xs.numel() * 4 + 2 * n_out)
report("bilinear_bwd", timeit(lambda: L.check(lib.gdn_bilinear_bwd(yb.data_ptr(), xs.data_ptr(), B, 64, 128, 256, 512, Cc, 0, S()))), xs.numel() * 4 + n_out)
xp = torch.randn(B, 256, 512, 64, device=dev)
yp = torch.empty(B, 128, 256, 64, device=dev)
report("maxpool2_fwd 256x512x64", timeit(lambda: L.check(lib.gdn_maxpool2_fwd(xp.data_ptr(), yp.data_ptr(), B, 256, 512, 64, S()))), xp.numel() * 4 + yp.numel() * 4)
dxp = torch.empty_like(xp)
report("maxpool2_bwd", timeit(lambda: L.check(lib.gdn_maxpool2_bwd(xp.data_ptr(), yp.data_ptr(), dxp.data_ptr(), B, 256, 512, 64, S()))), 2 * xp.numel() * 4 + yp.numel() * 4)
a, b = torch.randn(B, 256, 512, 64, device=dev), torch.randn(B, 256, 512, 64, device=dev)
loss, g = torch.zeros(1, device=dev), torch.empty_like(a)
wsd = E.dot_ws(dev)
report("l1 loss + grad 1 GB", timeit(lambda: L.check(lib.gdn_l1(a.data_ptr(), b.data_ptr(), a.numel(), loss.data_ptr(), 1, g.data_ptr(), 1.0, 0, wsd.data_ptr(), S()))), 3 * a.numel() * 4)
p = torch.randn(268_435_456 // 4, device=dev); gr = torch.randn_like(p); m = torch.zeros_like(p); v = torch.zeros_like(p)
report("adamw 67M params", timeit(lambda: L.check(lib.gdn_adamw(p.data_ptr(), gr.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), 1e-3, 0.5, 0.999, 1e-8, 1e-4, 1, 1.0, S()))), 28 * p.numel())
