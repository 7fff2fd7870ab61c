"""Timing of the generator head's pieces (engine.op_upsample_skip_final) at the bench shapes: narrow 1x1 kernels vs achieved GB/s."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gan_danet_b200 import _lib as L, engine as E
dev = torch.device("cuda:0")
lib = L.lib_for_device(0)
B, H, W, Cc, T = 32, 128, 256, 64, 12
M = B * H * W
u = torch.randn(M, Cc, device=dev); wp = torch.randn(T, Cc, device=dev); z = torch.empty(M, T, device=dev); du = torch.empty(M, Cc, device=dev); gwp = torch.empty(T, Cc, device=dev)
ws = torch.empty(lib.gdn_narrow_conv1x1_wgrad_ws_bytes(M, Cc, T), dtype=torch.uint8, device=dev)
Z = torch.empty(B, 2 * H, 2 * W, T, device=dev); zs = torch.randn(B, H // 2, W // 2, T, device=dev); y = torch.empty(B, 2 * H, 2 * W, device=dev)
dz = torch.empty(M, T, device=dev)
st = E._stream
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
cases = {
 "narrow fwd   (read u 268 MB, write z 50 MB)": (lambda: lib.gdn_narrow_conv1x1_fwd(u.data_ptr(), Cc, wp.data_ptr(), T, z.data_ptr(), M, st()), 318e6),
 "narrow dgrad (read dz 50 MB, write du 268 MB)": (lambda: lib.gdn_narrow_conv1x1_dgrad(z.data_ptr(), T, wp.data_ptr(), Cc, du.data_ptr(), M, 0, st()), 318e6),
 "narrow wgrad (read u 268 + dz 50 MB)": (lambda: lib.gdn_narrow_conv1x1_wgrad(z.data_ptr(), T, u.data_ptr(), Cc, M, gwp.data_ptr(), 0, ws.data_ptr(), ws.numel(), st()), 318e6),
 "up2 + bilinear add on 12 planes (write 201 MB)": (lambda: lib.gdn_bicubic_up2_bilinear_add_fwd(z.data_ptr(), zs.data_ptr(), Z.data_ptr(), B, H, W, H // 2, W // 2, T, st()), 252e6),
 "tap_shift_sum (read 201 MB)": (lambda: lib.gdn_tap_shift_sum(Z.data_ptr(), T, None, y.data_ptr(), B, 2 * H, 2 * W, st()), 218e6),
 "tap_shift_expand (write 201 MB)": (lambda: lib.gdn_tap_shift_expand(y.data_ptr(), Z.data_ptr(), T, B, 2 * H, 2 * W, st()), 218e6),
 "bicubic_up2_bwd on 12 planes": (lambda: lib.gdn_bicubic_up2_bwd(Z.data_ptr(), dz.data_ptr(), B, H, W, T, st()), 252e6),
 "bilinear_bwd on 12 planes": (lambda: lib.gdn_bilinear_bwd(Z.data_ptr(), zs.data_ptr(), B, H // 2, W // 2, 2 * H, 2 * W, T, 0, st()), 214e6),
}
for k, (fn, nbytes) in cases.items():
    ms = t(fn)
    print(f"{ms:7.3f} ms  {nbytes / ms / 1e6:7.0f} GB/s  {k}")
