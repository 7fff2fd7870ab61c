"""Achieved HBM bandwidth of the post-processing / ensemble-statistics kernels (ensemble.cu) on full-size product fields:
T = 181 monthly fields of 256x512 (BASELINE grid output), M = 8 members.  Algorithmic bytes = every input read once + every output
written once at 4 B per element (1 B per mask pixel); CUDA events, 3 warm-up + 10 timed launches, buffers (190 MB per field stack,
1.5 GB for the member stack) exceed the 126 MB L2.  Prints one JSON line with GB/s and the fraction of MEASURED_PEAKS.json hbm_gbs."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gan_danet_b200 import postprocess as PP

dev = torch.device("cuda:0")
peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
T, H, W, M = 181, 256, 512, 8
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(T, 1, H, W, device=dev, generator=g)
trend = torch.randn(T, 1, H, W, device=dev, generator=g)
keep = torch.rand(H, W, device=dev, generator=g) > 0.3
members = torch.randn(M, T, 1, H, W, device=dev, generator=g)
ref = torch.randn(T, 1, H, W, device=dev, generator=g)
small = torch.randn(T, 1, H // 4, W // 4, device=dev, generator=g)
n = T * H * W


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


s_sorted = torch.sort(x.reshape(T, -1), dim=1).values
r_sorted = torch.sort(ref.reshape(T, -1), dim=1).values
out = torch.empty_like(x).reshape(T, -1)
from gan_danet_b200 import _lib as L, engine as E
lib = E._lib(x)
xb = x.clone()
cases = {
    "destandardise(+trend, mask)": (lambda: PP.destandardise(x, 9.5, 1.25, trend, keep), 12 * n),
    "masked_spatial_mean": (lambda: PP.masked_spatial_mean(x, keep), 4 * n),
    "ensemble_stats(M=8, mean+std)": (lambda: PP.ensemble_stats(members), (4 * M + 8) * n),
    "hist_match kernel (sorted copies given)": (lambda: L.check(lib.gdn_hist_match(x.data_ptr(), s_sorted.data_ptr(), r_sorted.data_ptr(), out.data_ptr(), T, H * W, H * W, 0.2, E._stream())), 8 * n),
    "hist_match incl. both row sorts (gdn_sort_rows)": (lambda: PP.hist_match(x, ref, 0.2), 8 * n),
    "bicubic_resize x4 (64x128 -> 256x512)": (lambda: PP.bicubic_resize(small, 4), 4 * n + 4 * n // 16),
    "blend_region (in place)": (lambda: PP.smooth_blend(xb, ref, (8, H - 8, 8, W - 8)), 12 * n),
}
res = {}
for name, (fn, nbytes) in cases.items():
    ms = timeit(fn)
    res[name] = {"ms": ms, "algorithmic_GB": nbytes / 1e9, "GBps": nbytes / ms / 1e6, "frac_of_hbm_peak": nbytes / ms / 1e6 / peak}
print(json.dumps({"workload": f"T={T} fields of {H}x{W}, M={M} members", "hbm_peak_GBps": peak, "kernels": res}))
