"""Host-side engine: thin wrappers over the C ABI plus a small reverse-mode tape.

All activations are NHWC float32 views (channel slices of wider buffers are allowed: the kernels take a pitch).
The tape replaces PyTorch autograd *inside* a module so that buffers can be written in place (dense-block concat
slices, gradient accumulation into slices); the module as a whole is exposed to PyTorch autograd through
``TapeFunction`` (one ``torch.autograd.Function`` per top-level call), so ``.backward()`` and ``torch.optim`` work
unchanged and gradients land in the ordinary ``.grad`` fields.

PyTorch is used for memory, streams and autograd plumbing only; every arithmetic op on the path is a CUDA kernel of
``libgandanet_sm100.so``.  No CPU fallback exists.
"""
from __future__ import annotations

import ctypes as C
import os
import warnings
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib as L
from ._lib import ACT_LRELU, ACT_NONE, ACT_RELU, PREC_BF16, PREC_BF16X3, PREC_FP16, PREC_FP16X3, PREC_FP32

PAM_TC_PRECISIONS = (PREC_FP16, PREC_FP16X3)     # fused tcgen05 flash kernels: fp16 logit operands / fp16 hi+lo split logit operands
PAM_PRECISION_NAMES = {"fp32": PREC_FP32, "fp16": PREC_FP16, "fp16x3": PREC_FP16X3}

Tensor = torch.Tensor

# ----------------------------------------------------------------------------------------------------------------
# plumbing
# ----------------------------------------------------------------------------------------------------------------


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_get_device = getattr(torch._C, "_cuda_getDevice", None)


def _stream() -> C.c_void_p:
    """Raw handle of PyTorch's current stream on the current device.  torch.cuda.current_stream() builds a Python Stream object
    (~15 us; ~600 launches per step = 8 ms of the host's 48 ms): the raw-handle query is two plain C calls."""
    if _raw_stream is None or _get_device is None:
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)
    return C.c_void_p(_raw_stream(_get_device()))


def _lib(t: Tensor):
    """The library, initialised for the tensor's device.  Launches go to the CURRENT device's current stream (``_stream``), so a tensor that lives
    on another device would be handed to the wrong GPU as a foreign pointer: refuse it here (the module entry points switch devices themselves)."""
    if not t.is_cuda:
        raise L.GdnError("gan_danet_b200 kernels need CUDA tensors on an sm_100 device (no CPU fallback)")
    idx = t.device.index if t.device.index is not None else torch.cuda.current_device()
    cur = _get_device() if _get_device is not None else torch.cuda.current_device()
    if idx != cur:
        raise L.GdnError(f"tensor on cuda:{idx} but the current device is cuda:{cur}: wrap the call in `with torch.cuda.device({idx}):` "
                         "(TapeModule.forward and the trainers do)")
    return L.lib_for_device(idx)


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


_workspaces: Dict[Tuple[int, int, str], Tensor] = {}
pam_bwd_tensor_core: bool = os.environ.get("GDN_PAM_BWD", "tc").lower() != "fp32"   # fused tcgen05 backward when the forward ran on tensor cores
kernel_timing: Optional[dict] = None   # set to {} by bench.py: family -> [(start event, end event, algorithmic FLOPs)] on the launching stream


def workspace(name: str, nbytes: int, device) -> Tensor:
    """Scratch buffer per (device, stream, purpose): two streams of one device never share a workspace."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, _raw_stream(idx) if _raw_stream is not None else torch.cuda.current_stream(device).cuda_stream, name)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


_stat_counters: Dict[Tuple[int, int], Tensor] = {}
fused_stat_tail: bool = os.environ.get("GDN_FUSED_STATS", "1") != "0"   # single-launch BatchNorm statistics / backward reductions (gdn_bn_stats, gdn_bn_bwd_reduce_f)


def stat_counters(device) -> Tensor:
    """Ticket counters of the single-launch reductions: zero on entry, restored to zero by the kernel; one buffer per (device, stream)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, _raw_stream(idx) if _raw_stream is not None else torch.cuda.current_stream(device).cuda_stream)
    buf = _stat_counters.get(key)
    if buf is None:
        buf = _stat_counters[key] = torch.zeros(L.load().gdn_stat_fused_counters(), dtype=torch.int32, device=device)
    return buf


def _timed(family: str, flops: float, fn: Callable[[], None], detail: str = "") -> None:
    """Runs ``fn`` (one C-ABI call); when bench.py collects timings, brackets it with CUDA events on the current stream."""
    if kernel_timing is None:
        fn()
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    kernel_timing.setdefault(family, []).append((e0, e1, flops, detail))


def _f32(t: Tensor) -> None:
    if t.dtype != torch.float32:
        raise L.GdnError(f"expected float32 tensor, got {t.dtype}")


def pitch_of(t: Tensor) -> int:
    """Row pitch (in floats) of an NHWC-style view [..., C]; checks that rows are uniformly strided."""
    _f32(t)
    if t.shape[-1] > 1 and t.stride(-1) != 1:
        raise L.GdnError("innermost dimension must be contiguous")
    p = None
    expect = None
    for d in range(t.dim() - 2, -1, -1):
        if t.shape[d] == 1:
            continue
        if p is None:
            p = t.stride(d)
            expect = p * t.shape[d]
        else:
            if t.stride(d) != expect:
                raise L.GdnError(f"non-uniform row stride: shape {tuple(t.shape)} strides {t.stride()}")
            expect *= t.shape[d]
    if p is None:
        p = t.shape[-1]
    if p < t.shape[-1]:
        raise L.GdnError(f"overlapping rows: shape {tuple(t.shape)} strides {t.stride()}")
    return p


def rows_of(t: Tensor) -> int:
    n = 1
    for s in t.shape[:-1]:
        n *= s
    return n


def new_nhwc(B: int, H: int, W: int, Cc: int, like: Tensor) -> Tensor:
    return torch.empty((B, H, W, Cc), dtype=torch.float32, device=like.device)


# ----------------------------------------------------------------------------------------------------------------
# raw kernel wrappers
# ----------------------------------------------------------------------------------------------------------------


def fill_(t: Tensor, v: float) -> None:
    assert t.is_contiguous()
    L.check(_lib(t).gdn_fill(t.data_ptr(), t.numel(), float(v), _stream()), "gdn_fill")


def nchw_to_nhwc(src: Tensor, dst: Tensor) -> None:
    B, Cc, H, W = src.shape
    assert src.is_contiguous()
    L.check(_lib(src).gdn_nchw_to_nhwc(src.data_ptr(), dst.data_ptr(), pitch_of(dst), 0, B, Cc, H, W, _stream()), "gdn_nchw_to_nhwc")


def nhwc_to_nchw(src: Tensor, dst: Tensor) -> None:
    B, H, W, Cc = src.shape
    assert dst.is_contiguous()
    L.check(_lib(src).gdn_nhwc_to_nchw(src.data_ptr(), pitch_of(src), 0, dst.data_ptr(), B, Cc, H, W, _stream()), "gdn_nhwc_to_nchw")


def weight_ohwi(w: Tensor) -> Tensor:
    O, I, kh, kw = w.shape
    if kh == 1 and kw == 1:
        return w.detach().reshape(O, 1, 1, I)
    out = torch.empty((O, kh, kw, I), dtype=torch.float32, device=w.device)
    L.check(_lib(w).gdn_weight_oihw_to_ohwi(w.detach().contiguous().data_ptr(), out.data_ptr(), O, I, kh, kw, _stream()), "weight_ohwi")
    return out


def weight_ihwo(w: Tensor) -> Tensor:
    O, I, kh, kw = w.shape
    out = torch.empty((I, kh, kw, O), dtype=torch.float32, device=w.device)
    L.check(_lib(w).gdn_weight_oihw_to_ihwo(w.detach().contiguous().data_ptr(), out.data_ptr(), O, I, kh, kw, _stream()), "weight_ihwo")
    return out


def conv_raw(x: Tensor, w4: Tensor, y: Tensor, *, kh: int, kw: int, stride: int = 1, pad: int = 0, transposed: bool = False,
             bias: Optional[Tensor] = None, act: int = ACT_NONE, slope: float = 0.0, res: Optional[Tensor] = None,
             alpha: Optional[Tensor] = None, w_c0: int = 0, cin: Optional[int] = None) -> None:
    """y = act(alpha * conv(x, w) + bias) + res.  x: [B,Hi,Wi,Cin] view, w4: [Cout,kh,kw,Ctot] contiguous, y: [B,Ho,Wo,Cout] view."""
    lib = _lib(x)
    a = L.ConvArgs()
    B, Hi, Wi, Cx = x.shape
    _, Ho, Wo, Cout = y.shape
    Cin = Cx if cin is None else cin
    a.x, a.x_pitch, a.x_c0 = x.data_ptr(), pitch_of(x), 0
    a.w, a.w_k_pitch, a.w_c0, a.w_group_stride, a.groups = w4.data_ptr(), w4.shape[-1], w_c0, 0, 1
    a.y, a.y_pitch, a.y_c0 = y.data_ptr(), pitch_of(y), 0
    a.bias, a.alpha_ptr = _ptr(bias), _ptr(alpha)
    if res is not None:
        a.res, a.res_pitch, a.res_c0 = res.data_ptr(), pitch_of(res), 0
    a.B, a.Hi, a.Wi, a.Cin, a.Ho, a.Wo, a.Cout = B, Hi, Wi, Cin, Ho, Wo, Cout
    a.kh, a.kw, a.stride, a.pad, a.transposed = kh, kw, stride, pad, int(transposed)
    a.act, a.slope = act, slope
    a.splits = lib.gdn_conv2d_suggest_splits(C.byref(a))
    if a.splits > 1:
        need = a.splits * B * Ho * Wo * Cout * 4
        buf = workspace("conv", need, x.device)
        a.ws, a.ws_bytes = buf.data_ptr(), buf.numel()
    L.check(lib.gdn_conv2d(C.byref(a), _stream()), "gdn_conv2d")


def wgrad_raw(dy: Tensor, x: Tensor, out: Tensor, *, kh: int, kw: int, stride: int = 1, pad: int = 0, layout: int = 1,
              out_cin_total: Optional[int] = None, out_c0: int = 0, accumulate: bool = False, cin: Optional[int] = None) -> None:
    """layout 1: out is an OIHW weight gradient; layout 0: out[Cout][kh*kw*Cin]."""
    lib = _lib(x)
    a = L.WgradArgs()
    B, Hi, Wi, Cx = x.shape
    _, Ho, Wo, Cout = dy.shape
    Cin = Cx if cin is None else cin
    a.dy, a.dy_pitch, a.dy_c0 = dy.data_ptr(), pitch_of(dy), 0
    a.x, a.x_pitch, a.x_c0 = x.data_ptr(), pitch_of(x), 0
    a.out, a.layout, a.out_cin_total, a.out_c0, a.accumulate = out.data_ptr(), layout, (out_cin_total or Cin), out_c0, int(accumulate)
    a.scale_ptr, a.scale = None, 1.0
    a.B, a.Hi, a.Wi, a.Cin, a.Ho, a.Wo, a.Cout = B, Hi, Wi, Cin, Ho, Wo, Cout
    a.kh, a.kw, a.stride, a.pad, a.groups = kh, kw, stride, pad, 1
    a.splits = lib.gdn_wgrad_suggest_splits(C.byref(a))
    if a.splits > 1:
        need = a.splits * Cout * kh * kw * Cin * 4
        buf = workspace("wgrad", need, x.device)
        a.ws, a.ws_bytes = buf.data_ptr(), buf.numel()
    L.check(lib.gdn_conv2d_wgrad(C.byref(a), _stream()), "gdn_conv2d_wgrad")


# ----------------------------------------------------------------------------------------------------------------
# convolution precision: 'fp32' = CUDA-core engine (igemm_simt.cu); 'bf16' / 'bf16x3' = tcgen05 implicit GEMM (conv_tc.cu)
# ----------------------------------------------------------------------------------------------------------------

conv_precision: str = os.environ.get("GDN_CONV_PRECISION", "bf16x3").lower()
_PREC = {"bf16": PREC_BF16, "bf16x3": PREC_BF16X3}


def set_conv_precision(p: str) -> None:
    """'fp32' (CUDA-core parity engine), 'bf16' (tensor cores, bf16 operands) or 'bf16x3' (tensor cores, hi+lo split)."""
    global conv_precision
    p = p.lower()
    if p not in ("fp32", "bf16", "bf16x3"):
        raise ValueError(f"unknown conv precision {p!r}")
    conv_precision = p


class conv_precision_scope:
    """``with conv_precision_scope('bf16x3'):`` -- the convolutions RECORDED inside run their forward in that precision (their backward closures
    execute later, under whatever precision is current then).  Used for Discriminator1's forward in the product mode."""

    def __init__(self, p: Optional[str]):
        self.p = p

    def __enter__(self):
        global conv_precision
        self.old = conv_precision
        if self.p is not None:
            _outer_precision.append(conv_precision)
            set_conv_precision(self.p)

    def __exit__(self, *exc):
        global conv_precision
        if self.p is not None:
            _outer_precision.pop()
        conv_precision = self.old
        return False


_outer_precision: List[str] = []      # precision outside each active conv_precision_scope (innermost last)


def backward_conv_precision() -> str:
    """The precision the backward closures of what is being recorded will run under: the mode outside the outermost active conv_precision_scope."""
    return _outer_precision[0] if _outer_precision else conv_precision


def split_forward_in_product_mode() -> bool:
    """True while a forward is recorded with hi+lo split operands inside the product mode (generator_forward_x3, Discriminator1's forward).  Two
    operand shortcuts of the bf16 mode stay valid there because they do not touch what the split forward computes: the value projection's epilogue may
    still emit the fused PAM's bf16 V operand (the kernel rounds V to bf16 itself: the same round-to-nearest of the same fp32 value), and BatchNorm's
    backward may still hand the convolution its dz as a bf16 operand only (the gradient GEMMs run under 'bf16' and would round it anyway)."""
    return conv_precision == "bf16x3" and backward_conv_precision() == "bf16"


# Discriminator1 has no normalisation layer: the bf16 operand rounding of its three tensor-core convolutions reaches the logits directly, and
# loss_D = BCE(logits) moved by up to 1.5e-2 against the reference in round 1 (200-step teacher-forced walk).  In the product mode its FORWARD
# convolutions therefore use hi+lo split operands (3 MMA passes on 0.35 TF per step: +1 ms of 57); the backward stays bf16.
discriminator_forward_x3: bool = os.environ.get("GDN_D_FORWARD_X3", "1") != "0"


def discriminator_forward_precision() -> Optional[str]:
    return "bf16x3" if (discriminator_forward_x3 and conv_precision == "bf16") else None


# The generator's FORWARD convolutions on hi+lo split operands inside the product mode (off by default; bench.py and GANTrainer callers switch it on).
# bf16 operand rounding of the forward puts the generated field 9e-3 from the reference's (float64 run of generator.py:230-247), which is what moves
# loss_D / loss_G by more than 1 % on a few of 200 teacher-forced steps (tools/trajectory_quantised_cpu.py shows the same with the reference's own
# modules and bf16 operands: the format, not the kernels).  With split forward operands the field is 1.5e-4 away (the parity mode's forward) while
# every gradient GEMM still reads single bf16 operands (the hi parts): 3x the MMA work on the 2.4 ms of generator forward convolutions only
# (measured on B200, same box, graph replay: 54.98 vs 50.46 ms per step; the all-split parity mode: 91.6 ms).
# Quantised-oracle prediction (oracle/quantised_oracle.py Formats.forward_x3): y 1.5e-4, dx 4.1e-2, parameter gradients 3.8e-2 (bf16 forward: 9.0e-3 /
# 1.9e-1 / 1.8e-1 -- the sqrt law of DESIGN.md 4: no flipped ReLU masks, no wrong gradients).
generator_forward_x3: bool = os.environ.get("GDN_G_FORWARD_X3", "0") == "1"


def generator_forward_precision() -> Optional[str]:
    return "bf16x3" if (generator_forward_x3 and conv_precision == "bf16") else None


def tc_eligible(cin: int, cout: int, kh: int, kw: int, stride: int, ho: int, wo: int) -> bool:
    """Shapes the tensor-core kernel takes; the rest (fully connected layers, 1- and 3-channel inputs) stay on the
    fp32 CUDA-core engine, which is HBM-bound there anyway."""
    return (conv_precision in _PREC and kh == kw and kh in (1, 3) and stride in (1, 2) and cin >= 16 and ho * wo >= 32)


class Packed:
    """bf16 tensor-core operand: hi (and lo for the split precision), [rows, round_up(C, 8)]."""
    __slots__ = ("hi", "lo", "C")

    def __init__(self, hi: Tensor, lo: Optional[Tensor], Cc: int):
        self.hi, self.lo, self.C = hi, lo, Cc


def pack_act(x: Tensor, scale: Optional[Tensor] = None, shift: Optional[Tensor] = None, act: int = ACT_NONE, slope: float = 0.0) -> Packed:
    """fp32 NHWC view -> bf16 operand (optionally with a fused per-channel affine + activation, i.e. BatchNorm+ReLU)."""
    M, Cc = rows_of(x), x.shape[-1]
    Cp = (Cc + 7) // 8 * 8
    split = conv_precision == "bf16x3"
    hi = torch.empty((M, Cp), dtype=torch.bfloat16, device=x.device)
    lo = torch.empty((M, Cp), dtype=torch.bfloat16, device=x.device) if split else None
    L.check(_lib(x).gdn_pack_act_bf16(x.data_ptr(), pitch_of(x), 0, M, Cc, hi.data_ptr(), _ptr(lo), _ptr(scale), _ptr(shift), act, slope, _stream()),
            "gdn_pack_act_bf16")
    return Packed(hi, lo, Cc)


def pack_actgrad(dy: Tensor, y: Tensor, act: int, slope: float = 0.0) -> Packed:
    """bf16 operand of dz = dy * act'(y): the activation backward of a fused conv + ReLU/LeakyReLU done while packing, so that
    the fp32 dz never exists (12 -> 10 bytes per element less than act_bwd + pack_act, and one launch less)."""
    M, Cc = rows_of(dy), dy.shape[-1]
    Cp = (Cc + 7) // 8 * 8
    split = conv_precision == "bf16x3"
    hi = torch.empty((M, Cp), dtype=torch.bfloat16, device=dy.device)
    lo = torch.empty((M, Cp), dtype=torch.bfloat16, device=dy.device) if split else None
    L.check(_lib(dy).gdn_pack_actgrad_bf16(dy.data_ptr(), pitch_of(dy), y.data_ptr(), pitch_of(y), M, Cc, hi.data_ptr(), _ptr(lo), act, slope, _stream()),
            "gdn_pack_actgrad_bf16")
    return Packed(hi, lo, Cc)


def pack_actgrad16(dy: Tensor, y16: Tensor, act: int, slope: float = 0.0) -> Packed:
    """pack_actgrad with the activation output given as bf16 [M, Cp] (bf16-only feature maps of the frozen VGG19 branch)."""
    M, Cc = rows_of(dy), dy.shape[-1]
    Cp = (Cc + 7) // 8 * 8
    hi = torch.empty((M, Cp), dtype=torch.bfloat16, device=dy.device)
    L.check(_lib(dy).gdn_pack_actgrad_bf16g(dy.data_ptr(), pitch_of(dy), y16.data_ptr(), y16.shape[-1], M, Cc, hi.data_ptr(), None, act, slope, _stream()),
            "gdn_pack_actgrad_bf16g")
    return Packed(hi, None, Cc)


bf16_feature_storage: bool = os.environ.get("GDN_BF16_FEATURES", "1") != "0"


def bf16_storage_ok() -> bool:
    """bf16-only feature maps are used where the convolutions round their operands to bf16 anyway (conv precision 'bf16')."""
    return bf16_feature_storage and conv_precision == "bf16"


_frozen_weights: Dict[Tuple, Packed] = {}


def drop_frozen_weights(uid: int) -> None:
    """Evicts the packed weights of one PerceptualLoss instance (called by its weakref finalizer): keys start with (uid, ...)."""
    for k in [k for k in _frozen_weights if isinstance(k[0], tuple) and k[0] and k[0][0] == uid]:
        del _frozen_weights[k]


def pack_weight(w: Tensor, transposed: bool, frozen_key: Optional[Tuple] = None) -> Packed:
    """OIHW fp32 weight -> bf16 GEMM operand [taps][rows][K_p8].  ``frozen_key`` caches the result (VGG19 weights)."""
    split = conv_precision == "bf16x3"
    key = None
    if frozen_key is not None:
        key = (frozen_key, transposed, split, w.data_ptr(), w._version)
        hit = _frozen_weights.get(key)
        if hit is not None:
            return hit
    O, I, kh, kw = w.shape
    lib = _lib(w)
    n = lib.gdn_pack_weight_bf16_elems(O, I, kh, kw, int(transposed))
    hi = torch.empty(n, dtype=torch.bfloat16, device=w.device)
    lo = torch.empty(n, dtype=torch.bfloat16, device=w.device) if split else None
    wc = w.detach().contiguous()
    L.check(lib.gdn_pack_weight_bf16(wc.data_ptr(), O, I, 0, I, kh, kw, int(transposed), hi.data_ptr(), _ptr(lo), _stream()), "gdn_pack_weight_bf16")
    out = Packed(hi, lo, O if transposed else I)
    if key is not None:
        _frozen_weights[key] = out
    return out


def conv_tc_raw(xp: Packed, wp: Packed, y: Tensor, in_hw: Tuple[int, int], *, cin: int, kh: int, kw: int, stride: int = 1, pad: int = 0,
                transposed: bool = False, bias: Optional[Tensor] = None, act: int = ACT_NONE, slope: float = 0.0, res: Optional[Tensor] = None,
                y16: Optional[Tensor] = None, out_shape: Optional[Tuple[int, int, int, int]] = None) -> None:
    """Tensor-core convolution on packed operands.  ``in_hw`` is the grid of the packed input operand; y: [B,Ho,Wo,Cout] view.
    ``y16`` ([B*Ho*Wo, Cout] bf16, Cout % 8 == 0) additionally receives the output as the next convolution's packed operand;
    with ``y`` None (``out_shape`` given) the fp32 output is not written at all."""
    a = L.ConvTcArgs()
    B, Ho, Wo, Cout = y.shape if y is not None else out_shape
    a.x_hi, a.x_lo, a.w_hi, a.w_lo = xp.hi.data_ptr(), _ptr(xp.lo), wp.hi.data_ptr(), _ptr(wp.lo)
    if y is not None:
        a.y, a.y_pitch, a.y_c0 = y.data_ptr(), pitch_of(y), 0
    if y16 is not None:
        a.y16, a.y16_pitch = y16.data_ptr(), y16.shape[-1]
    a.bias = _ptr(bias)
    if res is not None:
        a.res, a.res_pitch, a.res_c0 = res.data_ptr(), pitch_of(res), 0
    a.B, a.Hi, a.Wi, a.Cin, a.Ho, a.Wo, a.Cout = B, in_hw[0], in_hw[1], cin, Ho, Wo, Cout
    a.kh, a.kw, a.stride, a.pad, a.transposed = kh, kw, stride, pad, int(transposed)
    a.act, a.slope, a.precision = act, slope, _PREC[conv_precision]
    _timed("conv_tc_fwd_kernel", 2.0 * B * Ho * Wo * Cout * cin * kh * kw / (stride * stride if transposed else 1),
           lambda: L.check(_lib(xp.hi).gdn_conv2d_tc(C.byref(a), _stream()), "gdn_conv2d_tc"),
           f"{'dgrad' if transposed else 'fwd'} B{B} {in_hw[0]}x{in_hw[1]}->{Ho}x{Wo} C{cin}->{Cout} k{kh} s{stride}")


def wgrad_tc_raw(dyp: Packed, xp: Packed, out: Tensor, *, B: int, in_hw: Tuple[int, int], out_hw: Tuple[int, int], cin: int, cout: int, kh: int, kw: int,
                 stride: int = 1, pad: int = 0, accumulate: bool = False) -> None:
    """Tensor-core weight gradient into an OIHW tensor."""
    lib = _lib(out)
    a = L.WgradTcArgs()
    a.dy_hi, a.dy_lo, a.x_hi, a.x_lo = dyp.hi.data_ptr(), _ptr(dyp.lo), xp.hi.data_ptr(), _ptr(xp.lo)
    a.out, a.out_cin_total, a.out_c0, a.accumulate, a.scale = out.data_ptr(), cin, 0, int(accumulate), 1.0
    a.B, a.Hi, a.Wi, a.Cin, a.Ho, a.Wo, a.Cout = B, in_hw[0], in_hw[1], cin, out_hw[0], out_hw[1], cout
    a.kh, a.kw, a.stride, a.pad, a.precision = kh, kw, stride, pad, _PREC[conv_precision]
    need = lib.gdn_conv2d_wgrad_tc_ws_bytes(C.byref(a))
    buf = workspace("wgrad_tc", need, out.device)
    a.ws, a.ws_bytes = buf.data_ptr(), buf.numel()
    _timed("conv_tc_wgrad_kernel", 2.0 * B * out_hw[0] * out_hw[1] * cout * cin * kh * kw,
           lambda: L.check(lib.gdn_conv2d_wgrad_tc(C.byref(a), _stream()), "gdn_conv2d_wgrad_tc"),
           f"wgrad B{B} {in_hw[0]}x{in_hw[1]}->{out_hw[0]}x{out_hw[1]} C{cin}->{cout} k{kh} s{stride}")


thin_conv_enabled: bool = os.environ.get("GDN_THIN_CONV", "1") != "0"
# relu1_1 of the perceptual loss recomputed from the single-channel images instead of stored (models/losses.py::_frozen_conv1_tap)
vgg_tap1_recompute: bool = os.environ.get("GDN_VGG_TAP1_RECOMPUTE", "1") != "0"


def _thin_kind(Cin: int, O: int, kh: int, kw: int, stride: int, x: Tensor, y: Tensor, act: int = ACT_NONE, bias=None, res=None) -> Optional[str]:
    """'expand' (1 -> C) / 'reduce' (C -> 1) when the HBM-bound thin-convolution kernels (thin_conv.cu) apply, else None."""
    if not thin_conv_enabled or kh != 3 or kw != 3:
        return None
    ok = L.load().gdn_thin_conv_supported
    if Cin == 1 and ok(O, kh, kw) and x.is_contiguous() and pitch_of(y) % 4 == 0 and y.data_ptr() % 16 == 0 \
            and (res is None or (pitch_of(res) % 4 == 0 and res.data_ptr() % 16 == 0)):
        return "expand"
    if O == 1 and stride == 1 and act == ACT_NONE and ok(Cin, kh, kw) and y.is_contiguous() and pitch_of(x) % 4 == 0 and x.data_ptr() % 16 == 0 \
            and (res is None or res.is_contiguous()):
        return "reduce"
    return None


class ConvCtx:
    """What a convolution keeps from its forward pass for the backward pass (the packed input on the tensor-core path)."""
    __slots__ = ("tc", "xp")

    def __init__(self, tc: bool, xp: Optional[Packed]):
        self.tc, self.xp = tc, xp


def conv_forward(x: Tensor, w: Tensor, y: Optional[Tensor], *, stride: int = 1, pad: int = 0, bias: Optional[Tensor] = None, act: int = ACT_NONE, slope: float = 0.0,
                 res: Optional[Tensor] = None, frozen_key: Optional[Tuple] = None, keep: bool = True, x_packed: Optional[Packed] = None,
                 y16: Optional[Tensor] = None) -> ConvCtx:
    """y = act(conv(x, w) + bias) + res for an OIHW weight; dispatches on ``conv_precision``.
    bf16 feature-map storage (frozen VGG19 branch): ``x_packed`` is the already packed input (then ``x`` only supplies the shape),
    ``y16`` receives the output as bf16 [B*Ho*Wo, O]; ``y`` may be None when only the bf16 copy is wanted (tensor-core path only)."""
    O, I, kh, kw = w.shape
    B, Hi, Wi, Cin = x.shape
    Ho, Wo = (Hi + 2 * pad - kh) // stride + 1, (Wi + 2 * pad - kw) // stride + 1
    if y is None or x_packed is not None:
        assert tc_eligible(Cin, O, kh, kw, stride, Ho, Wo) and (y16 is not None or y is not None)
        assert conv_precision == "bf16" or (y is not None and (y16 is None or split_forward_in_product_mode()))
        xp = x_packed if x_packed is not None else pack_act(x)
        conv_tc_raw(xp, pack_weight(w, False, frozen_key), y, (Hi, Wi), cin=Cin, kh=kh, kw=kw, stride=stride, pad=pad, bias=bias, act=act, slope=slope, res=res,
                    y16=y16, out_shape=(B, Ho, Wo, O))
        return ConvCtx(True, xp if keep else None)
    thin = _thin_kind(Cin, O, kh, kw, stride, x, y, act, bias, res)
    if thin == "expand":          # 1 -> C (Discriminator1.conv1, VGG19 conv1_1 on the channel-summed weight)
        L.check(_lib(x).gdn_thin_conv_expand_p(x.data_ptr(), w.contiguous().data_ptr(), _ptr(bias), y.data_ptr(), pitch_of(y), _ptr(res), pitch_of(res) if res is not None else 0,
                                               B, Ho, Wo, O, Hi, Wi, stride, pad, 0, act, slope, _ptr(y16), _stream()), "gdn_thin_conv_expand_p")
        return ConvCtx(False, None)
    if thin == "reduce":          # C -> 1 (the generator's final conv)
        L.check(_lib(x).gdn_thin_conv_reduce(x.data_ptr(), pitch_of(x), w.contiguous().data_ptr(), _ptr(bias), y.data_ptr(), _ptr(res),
                                             B, Hi, Wi, Cin, Ho, Wo, 1, pad, 0, _stream()), "gdn_thin_conv_reduce")
        return ConvCtx(False, None)
    if tc_eligible(Cin, O, kh, kw, stride, Ho, Wo):
        xp = pack_act(x)
        conv_tc_raw(xp, pack_weight(w, False, frozen_key), y, (Hi, Wi), cin=Cin, kh=kh, kw=kw, stride=stride, pad=pad, bias=bias, act=act, slope=slope, res=res,
                    y16=y16 if conv_precision == "bf16" else None)
        return ConvCtx(True, xp if keep else None)
    if frozen_key is not None:
        key = (frozen_key, "ohwi", w.data_ptr(), w._version)
        w4 = _frozen_weights.get(key)
        if w4 is None:
            w4 = _frozen_weights[key] = weight_ohwi(w)
    else:
        w4 = weight_ohwi(w)
    conv_raw(x, w4, y, kh=kh, kw=kw, stride=stride, pad=pad, bias=bias, act=act, slope=slope, res=res)
    return ConvCtx(False, None)


def conv_backward_tc_only(O: int, Cin: int, kh: int, kw: int, stride: int, Ho: int, Wo: int, Hi: int, Wi: int, need_gw: bool, need_gx: bool) -> bool:
    """True when every requested gradient of this convolution runs on the tensor-core path from a packed dz (no thin kernel, no
    fp32 engine): then the caller may pass ``dz_packed`` from pack_actgrad instead of an fp32 dz."""
    if conv_precision == "fp32" or (thin_conv_enabled and kh == 3 and kw == 3 and (Cin == 1 or O == 1)):
        return False
    ok_w = (not need_gw) or tc_eligible(max(O, 16), O, kh, kw, stride, Ho, Wo)
    ok_x = (not need_gx) or tc_eligible(O, Cin, kh, kw, stride, Hi, Wi)
    return ok_w and ok_x


def thin_gated_dgrad_ok(x: Tensor, w: Tensor, dy: Tensor, y: Tensor, stride: int, gx: Optional[Tensor], need_gw: bool, need_bias: bool) -> bool:
    """True when the data gradient of a 1 -> C convolution + activation can take act'(y) while loading dy (gdn_thin_conv_reduce_gated): nothing else
    needs the activation's input gradient (no weight / bias gradient requested: frozen VGG19 conv1_1, Discriminator1.conv1 in the generator step)."""
    O, I, kh, kw = w.shape
    return bool(not need_gw and not need_bias and gx is not None and gx.is_contiguous() and _thin_kind(I, O, kh, kw, stride, x, dy) == "expand"
                and pitch_of(y) % 4 == 0 and y.data_ptr() % 16 == 0)


def thin_gated_dgrad(dy: Tensor, y: Tensor, x: Tensor, w: Tensor, gx: Tensor, *, stride: int, pad: int, accumulate: bool, slope: float) -> None:
    """gx (+)= conv_transpose(dy * act'(y), w) for a 1 -> C convolution: the activation backward is fused into the gradient kernel's loads."""
    O = w.shape[0]
    B, Hi, Wi, _ = x.shape
    _, Ho, Wo, _ = dy.shape
    L.check(_lib(x).gdn_thin_conv_reduce_gated(dy.data_ptr(), pitch_of(dy), y.data_ptr(), pitch_of(y), float(slope), w.contiguous().data_ptr(), None, gx.data_ptr(),
                                               gx.data_ptr() if accumulate else None, B, Ho, Wo, O, Hi, Wi, stride, pad, 1, _stream()), "gdn_thin_conv_reduce_gated")


def conv_backward(ctx: ConvCtx, dz: Optional[Tensor], x: Tensor, w: Tensor, *, stride: int = 1, pad: int = 0, gw: Optional[Tensor] = None,
                  gx: Optional[Tensor] = None, gx_accumulate: bool = False, frozen_key: Optional[Tuple] = None, dz_packed: Optional[Packed] = None,
                  out_hw: Optional[Tuple[int, int]] = None) -> None:
    """Weight gradient into ``gw`` (OIHW, overwritten) and data gradient into ``gx`` (NHWC view; accumulated when asked).
    ``dz_packed`` (+ ``out_hw``) replaces the fp32 ``dz`` when conv_backward_tc_only() holds."""
    O, I, kh, kw = w.shape
    B, Hi, Wi, Cin = x.shape
    if dz_packed is not None:
        Ho, Wo = out_hw
        if gw is not None:
            xp = ctx.xp if ctx.xp is not None else pack_act(x)
            wgrad_tc_raw(dz_packed, xp, gw, B=B, in_hw=(Hi, Wi), out_hw=(Ho, Wo), cin=Cin, cout=O, kh=kh, kw=kw, stride=stride, pad=pad)
        if gx is not None:
            conv_tc_raw(dz_packed, pack_weight(w, True, frozen_key), gx, (Ho, Wo), cin=O, kh=kh, kw=kw, stride=stride, pad=pad, transposed=True,
                        res=gx if gx_accumulate else None)
        return
    _, Ho, Wo, _ = dz.shape
    thin = _thin_kind(Cin, O, kh, kw, stride, x, dz)
    if thin is not None and (gx is None or (gx.is_contiguous() if thin == "expand" else (pitch_of(gx) % 4 == 0 and gx.data_ptr() % 16 == 0))):
        lib = _lib(x)
        wc = w.contiguous()
        wide, field, C_w = (dz, x, O) if thin == "expand" else (x, dz, Cin)          # wide tensor V / single-channel field S
        Hv, Wv = wide.shape[1], wide.shape[2]
        Hs, Ws = field.shape[1], field.shape[2]
        if gw is not None:
            buf = workspace("thin_wgrad", lib.gdn_thin_conv_wgrad_ws_bytes(B, Hv, Wv, C_w), x.device)
            L.check(lib.gdn_thin_conv_wgrad(wide.data_ptr(), pitch_of(wide), field.data_ptr(), gw.data_ptr(), 0, B, Hv, Wv, C_w, Hs, Ws, stride if thin == "expand" else 1, pad,
                                            0 if thin == "expand" else 1, buf.data_ptr(), buf.numel(), _stream()), "gdn_thin_conv_wgrad")
        if gx is not None:
            if thin == "expand":      # data gradient of the 1 -> C conv: gather-reduce of dz into the single-channel input
                L.check(lib.gdn_thin_conv_reduce(dz.data_ptr(), pitch_of(dz), wc.data_ptr(), None, gx.data_ptr(), gx.data_ptr() if gx_accumulate else None,
                                                 B, Ho, Wo, O, Hi, Wi, stride, pad, 1, _stream()), "gdn_thin_conv_reduce")
            else:                     # data gradient of the C -> 1 conv: scatter of the single-channel dz through the flipped taps
                L.check(lib.gdn_thin_conv_expand(dz.data_ptr(), wc.data_ptr(), None, gx.data_ptr(), pitch_of(gx), gx.data_ptr() if gx_accumulate else None, pitch_of(gx),
                                                 B, Hi, Wi, Cin, Ho, Wo, 1, pad, 1, ACT_NONE, 0.0, _stream()), "gdn_thin_conv_expand")
        return
    # the backward GEMMs pick the tensor-core path on their own shapes: a 1-channel input (D conv1, VGG conv1_1) keeps its
    # forward on the CUDA cores but its gradients have 64-channel operands
    tc_w = gw is not None and tc_eligible(max(O, 16), O, kh, kw, stride, Ho, Wo)     # any Cout/Cin: narrow operands are zero-filled by TMA
    tc_x = gx is not None and tc_eligible(O, Cin, kh, kw, stride, Hi, Wi)
    dzp: Optional[Packed] = pack_act(dz) if (tc_w or tc_x) else None
    if gw is not None:
        if tc_w:
            xp = ctx.xp if ctx.xp is not None else pack_act(x)
            wgrad_tc_raw(dzp, xp, gw, B=B, in_hw=(Hi, Wi), out_hw=(Ho, Wo), cin=Cin, cout=O, kh=kh, kw=kw, stride=stride, pad=pad)
        else:
            wgrad_raw(dz, x, gw, kh=kh, kw=kw, stride=stride, pad=pad)
    if gx is not None:
        res = gx if gx_accumulate else None
        if tc_x:
            conv_tc_raw(dzp, pack_weight(w, True, frozen_key), gx, (Ho, Wo), cin=O, kh=kh, kw=kw, stride=stride, pad=pad, transposed=True, res=res)
        else:
            if frozen_key is not None:
                key = (frozen_key, "ihwo", w.data_ptr(), w._version)
                wt = _frozen_weights.get(key)
                if wt is None:
                    wt = _frozen_weights[key] = weight_ihwo(w)
            else:
                wt = weight_ihwo(w)
            conv_raw(dz, wt, gx, kh=kh, kw=kw, stride=stride, pad=pad, transposed=True, res=res)


def colstats(x: Tensor) -> Tensor:
    """double[2C]: per-channel sum and sum of squares over all rows of an NHWC view."""
    lib = _lib(x)
    M, Cc = rows_of(x), x.shape[-1]
    out = torch.empty(2 * Cc, dtype=torch.float64, device=x.device)
    buf = workspace("stat", lib.gdn_colstats_ws_bytes(M, Cc), x.device)
    L.check(lib.gdn_colstats(x.data_ptr(), pitch_of(x), 0, M, Cc, out.data_ptr(), buf.data_ptr(), _stream()), "gdn_colstats")
    return out


def colsum_f32(x: Tensor) -> Tensor:
    """float[C]: per-channel sum over all rows of an NHWC view (bias gradients, d gamma); double accumulation, ONE launch."""
    if not fused_stat_tail:
        return sums_to_float(colstats(x), x.shape[-1])
    lib = _lib(x)
    M, Cc = rows_of(x), x.shape[-1]
    out = torch.empty(2 * Cc, dtype=torch.float32, device=x.device)
    buf = workspace("stat", lib.gdn_stat_fused_ws_bytes(M, Cc), x.device)
    L.check(lib.gdn_colsums_f(x.data_ptr(), pitch_of(x), 0, M, Cc, None, out.data_ptr(), buf.data_ptr(), stat_counters(x.device).data_ptr(), _stream()), "gdn_colsums_f")
    return out[:Cc]


def sums_to_float(sums: Tensor, n: int, scale: float = 1.0) -> Tensor:
    out = torch.empty(n, dtype=torch.float32, device=sums.device)
    L.check(_lib(out).gdn_sums_to_float(sums.data_ptr(), out.data_ptr(), n, float(scale), _stream()), "gdn_sums_to_float")
    return out


def affine_act(x: Tensor, y: Tensor, scale: Tensor, shift: Tensor, act: int, slope: float = 0.0) -> None:
    L.check(_lib(x).gdn_affine_act(x.data_ptr(), pitch_of(x), 0, y.data_ptr(), pitch_of(y), 0, rows_of(x), x.shape[-1],
                                   scale.data_ptr(), shift.data_ptr(), act, slope, _stream()), "gdn_affine_act")


def act_bwd(dy: Tensor, y: Tensor, dz: Tensor, act: int, slope: float) -> None:
    L.check(_lib(dy).gdn_act_bwd(dy.data_ptr(), pitch_of(dy), 0, y.data_ptr(), pitch_of(y), 0, dz.data_ptr(), pitch_of(dz), 0,
                                 rows_of(dy), dy.shape[-1], act, slope, _stream()), "gdn_act_bwd")


def axpy(x: Tensor, y: Tensor, alpha: float = 1.0, accumulate: bool = True) -> None:
    L.check(_lib(x).gdn_axpy(x.data_ptr(), pitch_of(x), 0, y.data_ptr(), pitch_of(y), 0, rows_of(x), x.shape[-1], float(alpha),
                             int(accumulate), _stream()), "gdn_axpy")


def dot_ws(device) -> Tensor:
    return workspace("dot", L.load().gdn_dot_ws_bytes(0), device)


# ----------------------------------------------------------------------------------------------------------------
# tape
# ----------------------------------------------------------------------------------------------------------------


class Var:
    """A tensor on the tape plus its (lazily created) gradient.  ``parent`` marks a channel slice of a wider buffer."""
    __slots__ = ("t", "_g", "needs_grad", "parent", "c0", "c1", "packed", "g16", "grad16_only")

    def __init__(self, t: Tensor, needs_grad: bool = True, parent: Optional["Var"] = None, c0: int = 0, c1: int = 0):
        self.t, self._g, self.needs_grad, self.parent, self.c0, self.c1 = t, None, needs_grad, parent, c0, c1
        self.g16: Optional["Packed"] = None         # the gradient as a bf16 tensor-core operand only (op_conv_bn_act): no fp32 gradient tensor exists
        self.grad16_only = False                    # set on a convolution output whose single consumer (BatchNorm) may hand back g16 instead of g
        self.packed: Optional["Packed"] = None      # bf16 tensor-core operand of this value, consumed by op_conv instead of packing t again; after
        # op_bn_act(packed_only=True) it is the ONLY form in which the value exists (t is then a shape carrier)

    @property
    def g(self) -> Optional[Tensor]:
        """The gradient; a channel slice sees whatever has been accumulated into its parent's gradient buffer."""
        if self._g is None and self.parent is not None:
            pg = self.parent.g
            if pg is not None:
                self._g = pg[..., self.c0:self.c1]
        return self._g

    @g.setter
    def g(self, value: Optional[Tensor]) -> None:
        self._g = value

    def slice(self, c0: int, c1: int) -> "Var":
        return Var(self.t[..., c0:c1], self.needs_grad, self, c0, c1)

    def grad_target(self) -> Tuple[Tensor, bool]:
        """(buffer to write this Var's gradient into, whether the writer must accumulate)."""
        if self.g is not None:
            return self.g, True
        if self.parent is not None:
            pg, _ = self.parent.grad_target_zeroed()
            self.g = pg[..., self.c0:self.c1]
            return self.g, True
        self.g = torch.empty(self.t.shape, dtype=torch.float32, device=self.t.device)
        return self.g, False

    def grad_target_zeroed(self) -> Tuple[Tensor, bool]:
        if self.g is None:
            if self.parent is not None:
                return self.grad_target()
            self.g = torch.empty(self.t.shape, dtype=torch.float32, device=self.t.device)
            fill_(self.g, 0.0)
        return self.g, True

    def add_grad(self, g: Tensor) -> None:
        """Accumulate an already computed dense gradient tensor."""
        if self.g is None and self.parent is None:
            self.g = g
        else:
            tgt, _ = self.grad_target()
            if g.dim() >= 2:
                axpy(g, tgt, 1.0, True)
            else:
                axpy(g.reshape(1, -1), tgt.reshape(1, -1), 1.0, True)


class Tape:
    def __init__(self, record: bool = True):
        self.record = record
        self.ops: List[Callable[[], None]] = []
        self.consumed = False

    def push(self, fn: Callable[[], None]) -> None:
        if self.record:
            self.ops.append(fn)

    def backward(self) -> None:
        if self.consumed:
            raise L.GdnError("this module's backward has already run: its saved buffers were released (call forward again; retain_graph / "
                             "a second backward through the same output is not supported)")
        for fn in reversed(self.ops):
            fn()
        self.ops.clear()
        self.consumed = True


# ----------------------------------------------------------------------------------------------------------------
# differentiable ops
# ----------------------------------------------------------------------------------------------------------------


def op_conv(tape: Tape, x: Var, w: Var, bias: Optional[Var], *, stride: int = 1, pad: int = 0, act: int = ACT_NONE, slope: float = 0.0,
            out: Optional[Var] = None, y16: Optional[Tensor] = None) -> Var:
    """nn.Conv2d (+ fused bias and activation).  ``w`` holds the OIHW parameter."""
    O, I, kh, kw = w.t.shape
    B, Hi, Wi, Cin = x.t.shape
    assert Cin == I, f"conv: input has {Cin} channels, weight expects {I}"
    Ho = (Hi + 2 * pad - kh) // stride + 1
    Wo = (Wi + 2 * pad - kw) // stride + 1
    if out is None:
        out = Var(new_nhwc(B, Ho, Wo, O, x.t))
    cctx = conv_forward(x.t, w.t.detach(), out.t, stride=stride, pad=pad, bias=None if bias is None else bias.t.detach(), act=act, slope=slope,
                        keep=tape.record and w.needs_grad, x_packed=x.packed, y16=y16)
    y = out

    def bwd():
        if y.g16 is not None:          # BatchNorm backward produced dz directly as the gradient GEMMs' bf16 operand (op_conv_bn_act)
            assert y.g is None and act == ACT_NONE and bias is None
            gw16 = torch.empty_like(w.t) if w.needs_grad else None
            tgt16, acc16 = x.grad_target() if x.needs_grad else (None, False)
            conv_backward(cctx, None, x.t, w.t.detach(), stride=stride, pad=pad, gw=gw16, gx=tgt16, gx_accumulate=acc16, dz_packed=y.g16, out_hw=(Ho, Wo))
            if gw16 is not None:
                w.add_grad(gw16)
            return
        if y.g is None:
            return
        dy = y.g
        gw = torch.empty_like(w.t) if w.needs_grad else None
        tgt, acc = x.grad_target() if x.needs_grad else (None, False)
        need_bias = bias is not None and bias.needs_grad
        if act != ACT_NONE and not need_bias and conv_backward_tc_only(O, Cin, kh, kw, stride, Ho, Wo, Hi, Wi, gw is not None, tgt is not None):
            # activation backward fused into the packing of the gradient GEMMs' operand: no fp32 dz
            conv_backward(cctx, None, x.t, w.t.detach(), stride=stride, pad=pad, gw=gw, gx=tgt, gx_accumulate=acc,
                          dz_packed=pack_actgrad(dy, y.t, act, slope), out_hw=(Ho, Wo))
            if gw is not None:
                w.add_grad(gw)
            return
        if act in (ACT_RELU, ACT_LRELU) and thin_gated_dgrad_ok(x.t, w.t, dy, y.t, stride, tgt, gw is not None, need_bias):
            thin_gated_dgrad(dy, y.t, x.t, w.t.detach(), tgt, stride=stride, pad=pad, accumulate=acc, slope=slope if act == ACT_LRELU else 0.0)
            return
        if act != ACT_NONE:
            dz = torch.empty(y.t.shape, dtype=torch.float32, device=dy.device)
            act_bwd(dy, y.t, dz, act, slope)
        else:
            dz = dy
        if need_bias:
            bias.add_grad(colsum_f32(dz))
        conv_backward(cctx, dz, x.t, w.t.detach(), stride=stride, pad=pad, gw=gw, gx=tgt, gx_accumulate=acc)
        if gw is not None:
            w.add_grad(gw)

    tape.push(bwd)
    return y


linear_tensor_core: Optional[bool] = None     # None: TF32 tensor cores in the 'bf16' conv precision mode only; True / False force it


def _linear_tc(lib, Bn: int, Fout: int, Fin: int) -> bool:
    use = linear_tensor_core if linear_tensor_core is not None else conv_precision == "bf16"
    return bool(use) and bool(lib.gdn_linear_tc_supported(Bn, Fout, Fin))


def op_linear(tape: Tape, x: Var, w: Var, bias: Optional[Var], *, act: int = ACT_NONE, slope: float = 0.0) -> Var:
    """nn.Linear on a [B, F] matrix (+ fused activation).  Large layers (Discriminator1.fc1) stream their fp32 weight through the
    TF32 tcgen05 kernels of linear_tc.cu; the rest runs on the fp32 implicit-GEMM engine as a 1x1 convolution over a B x 1 x 1
    grid (the data gradient uses the reduction-form kernel so the [out,in] weight is never transposed)."""
    Bn, Fin = x.t.shape
    Fout = w.t.shape[0]
    lib = _lib(x.t)
    tc = _linear_tc(lib, Bn, Fout, Fin) and x.t.is_contiguous() and w.t.is_contiguous()
    x4 = x.t.view(Bn, 1, 1, Fin)
    y = Var(torch.empty((Bn, Fout), dtype=torch.float32, device=x.t.device))
    if tc:
        buf = workspace("linear_tc", lib.gdn_linear_tc_fwd_ws_bytes(Bn, Fout, Fin), x.t.device)
        _timed("linear_tc_kernel", 2.0 * Bn * Fout * Fin,
               lambda: L.check(lib.gdn_linear_tc_fwd(x.t.data_ptr(), w.t.data_ptr(), None if bias is None else bias.t.data_ptr(), y.t.data_ptr(), Bn, Fout, Fin, act, slope,
                                                     buf.data_ptr(), buf.numel(), _stream()), "gdn_linear_tc_fwd"), f"fwd {Bn}x{Fin}->{Fout}")
    else:
        conv_raw(x4, w.t.detach().view(Fout, 1, 1, Fin), y.t.view(Bn, 1, 1, Fout), kh=1, kw=1,
                 bias=None if bias is None else bias.t.detach(), act=act, slope=slope)

    def bwd():
        if y.g is None:
            return
        dy = y.g
        if act != ACT_NONE:
            dz = torch.empty_like(dy)
            act_bwd(dy, y.t, dz, act, slope)
        else:
            dz = dy
        dz = dz.contiguous()
        dz4 = dz.view(Bn, 1, 1, Fout)
        if bias is not None and bias.needs_grad:
            bias.add_grad(colsum_f32(dz))
        if w.needs_grad:
            gw = torch.empty_like(w.t)
            if tc and Bn % 32 == 0 and Fout % 256 == 0:
                wbuf = workspace("linear_tc_w", lib.gdn_linear_tc_wgrad_ws_bytes(Bn, Fout, Fin), dz.device)
                _timed("linear_tc_kernel", 2.0 * Bn * Fout * Fin,
                       lambda: L.check(lib.gdn_linear_tc_wgrad(dz.data_ptr(), x.t.data_ptr(), gw.data_ptr(), Bn, Fout, Fin, wbuf.data_ptr(), wbuf.numel(), _stream()),
                                       "gdn_linear_tc_wgrad"), f"wgrad {Bn}x{Fin}->{Fout}")
            else:
                wgrad_raw(dz4, x4, gw.view(Fout, Fin, 1, 1), kh=1, kw=1)
            w.add_grad(gw)
        if x.needs_grad:
            gx = torch.empty((Bn, Fin), dtype=torch.float32, device=dz.device)
            if tc:
                _timed("linear_tc_kernel", 2.0 * Bn * Fout * Fin,
                       lambda: L.check(lib.gdn_linear_tc_dgrad(dz.data_ptr(), w.t.data_ptr(), gx.data_ptr(), Bn, Fout, Fin, _stream()), "gdn_linear_tc_dgrad"), f"dgrad {Bn}x{Fin}->{Fout}")
            else:
                # dx[b][j] = sum_o dz[b][o] W[o][j]: "pixels" = output features o, dy' = dz^T [Fout][Bn], x' = W [Fout][Fin]
                dzt = torch.empty((Fout, Bn), dtype=torch.float32, device=dz.device)
                nhwc_to_nchw(dz.view(1, Bn, 1, Fout), dzt.view(1, Fout, Bn, 1))
                wgrad_raw(dzt.view(1, Fout, 1, Bn), w.t.detach().view(1, Fout, 1, Fin), gx, kh=1, kw=1, layout=0)
            x.add_grad(gx)

    tape.push(bwd)
    return y


class BNState:
    """Parameters and buffers of one nn.BatchNorm2d."""
    __slots__ = ("weight", "bias", "running_mean", "running_var", "num_batches_tracked", "eps", "momentum")

    def __init__(self, weight: Var, bias: Var, running_mean: Tensor, running_var: Tensor, num_batches_tracked: Optional[Tensor],
                 eps: float = 1e-5, momentum: float = 0.1):
        self.weight, self.bias, self.running_mean, self.running_var = weight, bias, running_mean, running_var
        self.num_batches_tracked, self.eps, self.momentum = num_batches_tracked, eps, momentum


conv_bn_packed_grad: bool = os.environ.get("GDN_CONV_BN_GRAD16", "1") != "0"


def op_conv_bn_act(tape: Tape, x: Var, w: Var, bn: BNState, *, training: bool, act: int = ACT_RELU, slope: float = 0.0, stride: int = 1, pad: int = 0,
                   out: Optional[Var] = None) -> Var:
    """act(BN(conv(x))) with a bias-free convolution (generator.py:187-190 initial, :147-150 fuse, :218-224 upsample).  In the bf16 product
    mode the gradient between BatchNorm and the convolution is never stored in fp32: BatchNorm's backward writes it as the bf16 operand of
    the convolution's weight- and data-gradient GEMMs (gdn_bn_bwd_apply16), which is the only form in which it is read."""
    z = op_conv(tape, x, w, None, stride=stride, pad=pad)
    O, Cin, kh, kw = w.t.shape
    _, Hi, Wi, _ = x.t.shape
    _, Ho, Wo, _ = z.t.shape
    z.grad16_only = bool(conv_bn_packed_grad and training and tape.record and (conv_precision == "bf16" or split_forward_in_product_mode()) and O % 8 == 0 and pitch_of(z.t) % 4 == 0
                         and z.t.data_ptr() % 16 == 0
                         and conv_backward_tc_only(O, Cin, kh, kw, stride, Ho, Wo, Hi, Wi, w.needs_grad, x.needs_grad))
    return op_bn_act(tape, z, bn, training=training, act=act, slope=slope, out=out)


fuse_bn_into_pack: bool = os.environ.get("GDN_FUSE_BN_PACK", "1") != "0"


def op_bn_act_conv(tape: Tape, x: Var, bn: BNState, w: Var, bias: Optional[Var], *, training: bool, act: int = ACT_RELU, slope: float = 0.0,
                   stride: int = 1, pad: int = 0, out: Optional[Var] = None) -> Var:
    """conv(act(BN(x))) (DenseLayer generator.py:32-34, TransitionLayer :61-63).  On the tensor-core path the normalised activation is
    never stored in fp32: the per-channel affine + ReLU is applied while the convolution's bf16 operand is packed (x read once, 2 B/element
    written, instead of 4 + 4 for the BN output and 4 + 2 for its packing); the backward needs x, the coefficients and the packed operand only."""
    O, _, kh, kw = w.t.shape
    _, Hi, Wi, Cin = x.t.shape
    Ho, Wo = (Hi + 2 * pad - kh) // stride + 1, (Wi + 2 * pad - kw) // stride + 1
    fused = (fuse_bn_into_pack and tc_eligible(Cin, O, kh, kw, stride, Ho, Wo) and not (thin_conv_enabled and kh == 3 and (Cin == 1 or O == 1)))
    h = op_bn_act(tape, x, bn, training=training, act=act, slope=slope, packed_only=fused)
    return op_conv(tape, h, w, bias, stride=stride, pad=pad, out=out)


def op_bn_act(tape: Tape, x: Var, bn: BNState, *, training: bool, act: int = ACT_RELU, slope: float = 0.0, out: Optional[Var] = None,
              update_running: bool = True, packed_only: bool = False) -> Var:
    """nn.BatchNorm2d (train: batch statistics + running-stat update; eval: running statistics) followed by an activation.
    ``packed_only``: the result is produced as the bf16 operand of the one tensor-core convolution that consumes it (``Var.packed``);
    the returned Var's tensor is x's own (shape carrier, never read)."""
    lib = _lib(x.t)
    Cc = x.t.shape[-1]
    M = rows_of(x.t)
    dev = x.t.device
    coef = torch.empty((4, Cc), dtype=torch.float32, device=dev)   # mean, invstd, scale, shift
    mean, invstd, scale, shift = coef[0], coef[1], coef[2], coef[3]
    if training:
        upd = update_running and bn.running_mean is not None
        nbt = bn.num_batches_tracked if upd else None
        if fused_stat_tail and (nbt is None or (nbt.dtype == torch.int64 and nbt.device == dev)):
            # statistics, finalize and the num_batches_tracked increment in ONE launch (the reduction's last block finishes it)
            buf = workspace("stat", lib.gdn_stat_fused_ws_bytes(M, Cc), dev)
            L.check(lib.gdn_bn_stats(x.t.data_ptr(), pitch_of(x.t), 0, M, Cc, bn.weight.t.data_ptr(), bn.bias.t.data_ptr(), bn.eps, bn.momentum,
                                     bn.running_mean.data_ptr() if upd else None, bn.running_var.data_ptr() if upd else None, _ptr(nbt),
                                     mean.data_ptr(), invstd.data_ptr(), scale.data_ptr(), shift.data_ptr(), buf.data_ptr(), stat_counters(dev).data_ptr(),
                                     _stream()), "gdn_bn_stats")
        else:
            sums = colstats(x.t)
            L.check(lib.gdn_bn_finalize(sums.data_ptr(), M, Cc, bn.weight.t.data_ptr(), bn.bias.t.data_ptr(), bn.eps, bn.momentum,
                                        bn.running_mean.data_ptr() if upd else None, bn.running_var.data_ptr() if upd else None,
                                        mean.data_ptr(), invstd.data_ptr(), scale.data_ptr(), shift.data_ptr(), _stream()), "gdn_bn_finalize")
            if nbt is not None:
                nbt.add_(1)
    else:
        L.check(lib.gdn_bn_eval_coeffs(bn.weight.t.data_ptr(), bn.bias.t.data_ptr(), bn.running_mean.data_ptr(), bn.running_var.data_ptr(),
                                       bn.eps, Cc, scale.data_ptr(), shift.data_ptr(), _stream()), "gdn_bn_eval_coeffs")
    if packed_only:
        assert out is None
        out = Var(x.t)
        out.packed = pack_act(x.t, scale, shift, act, slope)
    else:
        if out is None:
            out = Var(torch.empty(x.t.shape, dtype=torch.float32, device=dev))
        affine_act(x.t, out.t, scale, shift, act, slope)
    y = out

    def bwd():
        if y.g is None:
            return
        if not training:
            raise L.GdnError("backward through eval-mode BatchNorm is not supported")
        dy = y.g
        sums = torch.empty(2 * Cc, dtype=torch.float64, device=dev)
        gb = None
        if fused_stat_tail:
            gb = torch.empty(2 * Cc, dtype=torch.float32, device=dev)
            buf = workspace("stat", lib.gdn_stat_fused_ws_bytes(M, Cc), dev)
            L.check(lib.gdn_bn_bwd_reduce_f(dy.data_ptr(), pitch_of(dy), 0, x.t.data_ptr(), pitch_of(x.t), 0, M, Cc, mean.data_ptr(), invstd.data_ptr(),
                                            scale.data_ptr(), shift.data_ptr(), act, slope, sums.data_ptr(), gb.data_ptr(), buf.data_ptr(),
                                            stat_counters(dev).data_ptr(), _stream()), "gdn_bn_bwd_reduce_f")
        else:
            buf = workspace("stat", lib.gdn_colstats_ws_bytes(M, Cc), dev)
            L.check(lib.gdn_bn_bwd_reduce(dy.data_ptr(), pitch_of(dy), 0, x.t.data_ptr(), pitch_of(x.t), 0, M, Cc, mean.data_ptr(), invstd.data_ptr(),
                                          scale.data_ptr(), shift.data_ptr(), act, slope, sums.data_ptr(), buf.data_ptr(), _stream()), "gdn_bn_bwd_reduce")
        if x.needs_grad and x.grad16_only and x.g is None and x.parent is None and pitch_of(dy) % 4 == 0 and dy.data_ptr() % 16 == 0:
            g16 = torch.empty((M, Cc), dtype=torch.bfloat16, device=dev)
            L.check(lib.gdn_bn_bwd_apply16(dy.data_ptr(), pitch_of(dy), 0, x.t.data_ptr(), pitch_of(x.t), 0, g16.data_ptr(), Cc, M, Cc, mean.data_ptr(),
                                           invstd.data_ptr(), bn.weight.t.data_ptr(), scale.data_ptr(), shift.data_ptr(), act, slope, sums.data_ptr(),
                                           _stream()), "gdn_bn_bwd_apply16")
            x.g16 = Packed(g16, None, Cc)
        elif x.needs_grad:
            tgt, acc = x.grad_target()
            L.check(lib.gdn_bn_bwd_apply(dy.data_ptr(), pitch_of(dy), 0, x.t.data_ptr(), pitch_of(x.t), 0, tgt.data_ptr(), pitch_of(tgt), 0, int(acc),
                                         M, Cc, mean.data_ptr(), invstd.data_ptr(), bn.weight.t.data_ptr(), scale.data_ptr(), shift.data_ptr(),
                                         act, slope, sums.data_ptr(), None, None, _stream()), "gdn_bn_bwd_apply")
        if gb is None:
            gb = sums_to_float(sums, 2 * Cc)          # (sum g, sum g*xhat) = (dbias, dweight)
        gwb = (gb[Cc:], gb[:Cc])
        if bn.weight.needs_grad:
            bn.weight.add_grad(gwb[0])
        if bn.bias.needs_grad:
            bn.bias.add_grad(gwb[1])

    tape.push(bwd)
    return y


_pam_v16: Dict[Tuple[int, int, int], Tensor] = {}
pam_v16_from_conv: bool = os.environ.get("GDN_PAM_V16", "1") != "0"


def pam_v16_buffer(x: Tensor) -> Optional[Tensor]:
    """bf16 [B*N, 192] operand buffer for the fused PAM forward, or None when the value projection cannot emit it.  The value
    projection's tensor-core epilogue writes columns [0, C) (conv y16 output, pitch 192); column C = 1 (the softmax denominator comes
    out of the P.V product) and the zero tail are written ONCE here -- one persistent buffer per (device, rows, C), so the packing pass
    of gdn_pam_fwd only touches q and k (V is 80 % of its bytes)."""
    B, H, W, Cc = x.shape
    N = H * W
    if not (pam_v16_from_conv and (conv_precision == "bf16" or split_forward_in_product_mode()) and N % 128 == 0 and Cc < 192 and Cc % 4 == 0 and tc_eligible(Cc, Cc, 1, 1, 1, H, W)):
        return None
    key = (x.device.index or 0, B * N, Cc)
    buf = _pam_v16.get(key)
    if buf is None:
        buf = torch.zeros((B * N, 192), dtype=torch.bfloat16, device=x.device)
        buf[:, Cc] = 1.0
        _pam_v16[key] = buf
    return buf


def release_buffers() -> None:
    """Drops the engine's persistent device buffers (workspaces, packed frozen weights, PAM value-operand buffers); they are
    re-created on demand.  Do not call while a captured CUDA graph that baked their addresses is still in use."""
    _workspaces.clear()
    _stat_counters.clear()
    _pam_v16.clear()
    _frozen_weights.clear()


# Measured on B200 (same box, back to back): 58.4 -> 57.5 ms/step, fused PAM forward 833 -> 854 TFLOP/s inside the step.  The bf16 store is its own
# instantiation of the PAM kernel: as a run-time branch of the one kernel it cost the fp32-store launches 4 %.
danet_cat16: bool = os.environ.get("GDN_DANET_CAT16", "1") != "0"


def danet_cat16_ok(x: Tensor, pam_fp16: bool, fuse_cout: int) -> bool:
    """True when PAM and CAM can write their outputs ONLY as bf16 column blocks of the fuse convolution's packed operand (generator.py:156-157
    cat -> conv): product mode, fused tensor-core PAM on an aligned grid, tensor-core CAM, tensor-core fuse convolution."""
    B, H, W, Cc = x.shape
    return bool(danet_cat16 and pam_fp16 and conv_precision == "bf16" and (H * W) % 128 == 0 and Cc < 192 and Cc % 4 == 0 and _cam_tc(Cc)
                and tc_eligible(2 * Cc, fuse_cout, 3, 3, 1, H, W))


def op_pam_core(tape: Tape, x: Var, q: Var, k: Var, v: Var, gamma: Var, *, precision: int = PREC_FP32, out: Optional[Var] = None,
                v16: Optional[Tensor] = None, y16: Optional[Tensor] = None) -> Var:
    """Position-attention core: y = gamma * softmax(q k^T) v + x  (generator.py:115-122).  ``v16``: the value operand already packed
    by the value projection's epilogue (pam_v16_buffer)."""
    lib = _lib(x.t)
    B, H, W, Cc = x.t.shape
    N, d = H * W, q.t.shape[-1]
    dev = x.t.device
    tcp = precision in PAM_TC_PRECISIONS
    if tcp and N % 128 != 0 and pam_pad_to_tiles and d < 32 and Cc < 192 and Cc % 4 == 0:
        assert y16 is None
        return _op_pam_core_padded(tape, x, q, k, v, gamma, out, precision=precision)
    if tcp and (N % 128 != 0 or d > 32 or Cc >= 192 or Cc % 4 != 0):
        # shape outside the tensor-core kernels' tiling (the reference's channel counts 160/176/184 and d = 20..23 are inside): the fp32 CUDA-core
        # engine materialises N x N chunks and is ~15x slower -- say so instead of silently falling off the cliff
        if not pam_allow_fp32_fallback:
            raise L.GdnError(f"op_pam_core: shape N={N} d={d} C={Cc} is outside the fused tensor-core kernel (needs d <= 32, C < 192, C % 4 == 0); "
                             "pass precision 'fp32' explicitly or set engine.pam_allow_fp32_fallback = True")
        warnings.warn(f"PAM shape N={N} d={d} C={Cc} runs on the fp32 CUDA-core engine (about 15x slower than the fused tcgen05 kernel)", RuntimeWarning, stacklevel=2)
        precision, tcp = PREC_FP32, False
    if out is None:
        out = Var(torch.empty(x.t.shape, dtype=torch.float32, device=dev))
    o = torch.empty((B, N, Cc), dtype=torch.float32, device=dev)
    lse = torch.empty((B, N), dtype=torch.float32, device=dev)
    a = L.PamFwdArgs()
    a.q, a.k, a.qk_pitch, a.d = q.t.data_ptr(), k.t.data_ptr(), pitch_of(q.t), d
    assert pitch_of(k.t) == a.qk_pitch
    a.v, a.v_pitch = v.t.data_ptr(), pitch_of(v.t)
    a.x, a.x_pitch, a.gamma = x.t.data_ptr(), pitch_of(x.t), gamma.t.data_ptr()
    a.o, a.y, a.y_pitch, a.lse = o.data_ptr(), out.t.data_ptr(), pitch_of(out.t), lse.data_ptr()
    a.B, a.N, a.C, a.precision, a.chunk = B, N, Cc, precision, 0
    need = lib.gdn_pam_fwd_ws_bytes(C.byref(a))
    buf = workspace("pam", need, dev)
    a.ws, a.ws_bytes = buf.data_ptr(), buf.numel()
    if v16 is not None and tcp:
        assert v16.shape == (B * N, 192) and v16.dtype == torch.bfloat16
        a.v16 = v16.data_ptr()
    if y16 is not None:     # y only as a bf16 column block (a strided [B*N, C] view) of the consumer's packed operand: out.t is never written
        assert tcp and y16.dtype == torch.bfloat16 and y16.shape == (B * N, Cc) and y16.stride(1) == 1
        a.y, a.y16, a.y16_pitch = None, y16.data_ptr(), y16.stride(0)
    fam = "pam_flash_fwd_kernel" if tcp else "pam_fwd_fp32"
    _timed(fam, 2.0 * B * N * N * (d + Cc), lambda: L.check(lib.gdn_pam_fwd(C.byref(a), _stream()), "gdn_pam_fwd"))
    y = out

    def bwd():
        if y.g is None:
            return
        dy = y.g
        dq = torch.empty((B, H, W, d), dtype=torch.float32, device=dev)
        dk = torch.empty((B, H, W, d), dtype=torch.float32, device=dev)
        dv = torch.empty((B, H, W, Cc), dtype=torch.float32, device=dev)
        rowdot = torch.empty((B * N, 1), dtype=torch.float32, device=dev)
        b = L.PamBwdArgs()
        b.q, b.k, b.qk_pitch, b.d = q.t.data_ptr(), k.t.data_ptr(), pitch_of(q.t), d
        b.v, b.v_pitch = v.t.data_ptr(), pitch_of(v.t)
        b.o, b.lse, b.gamma = o.data_ptr(), lse.data_ptr(), gamma.t.data_ptr()
        b.dy, b.dy_pitch = dy.data_ptr(), pitch_of(dy)
        b.dq, b.dk, b.dv, b.rowdot = dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), rowdot.data_ptr()
        b.B, b.N, b.C, b.precision, b.chunk = B, N, Cc, (precision if pam_bwd_tensor_core else PREC_FP32), 0
        need_b = lib.gdn_pam_bwd_ws_bytes(C.byref(b))
        wsb = workspace("pam", need_b, dev)
        b.ws, b.ws_bytes = wsb.data_ptr(), wsb.numel()
        famb = "pam_flash_bwd_kernel" if b.precision in PAM_TC_PRECISIONS else "pam_bwd_fp32"
        _timed(famb, 4.0 * B * N * N * (d + Cc), lambda: L.check(lib.gdn_pam_bwd(C.byref(b), _stream()), "gdn_pam_bwd"))
        if gamma.needs_grad:
            gamma.add_grad(colsum_f32(rowdot))
        if q.needs_grad:
            q.add_grad(dq)
        if k.needs_grad:
            k.add_grad(dk)
        if v.needs_grad:
            v.add_grad(dv)
        if x.needs_grad:
            _accumulate(x, dy)

    tape.push(bwd)
    return y


pam_allow_fp32_fallback: bool = os.environ.get("GDN_PAM_FP32_FALLBACK", "0") == "1"    # shapes outside the fused kernel: raise (default) or warn + fp32 engine
pam_pad_to_tiles: bool = os.environ.get("GDN_PAM_PAD", "1") != "0"   # N % 128 != 0 (the authors' 45x22 grid, N = 990) on the tensor-core kernels
PAM_KEY_MASK = -30000.0     # exactly representable in fp16; exp2((q.k + PAM_KEY_MASK) * log2e - m) underflows to 0 (2^-125 on the polynomial lanes)


def _op_pam_core_padded(tape: Tape, x: Var, q: Var, k: Var, v: Var, gamma: Var, out: Optional[Var], pad_to: Optional[int] = None,
                        precision: int = PREC_FP16X3) -> Var:
    """The fused tcgen05 kernels tile N in blocks of 128.  Other grids are padded to the next multiple per sample and the
    padded KEYS are masked through a spare column of the zero-padded logit operands (d < 32): q gets a column of ones, k a column
    that is 0 on real positions and PAM_KEY_MASK on padded ones, so every padded logit is S = -30000 and its softmax weight is 0.
    Padded QUERY rows are computed and dropped; in the backward their cotangent is zero, so they contribute nothing to dK / dV."""
    B, H, W, Cc = x.t.shape
    N, d = H * W, q.t.shape[-1]
    Np = (N + 127) // 128 * 128 if pad_to is None else int(pad_to)      # pad_to: tests force padding on an already aligned grid
    assert Np % 128 == 0 and Np >= N
    dev = x.t.device

    def padded(t: Tensor, width: int) -> Tensor:
        buf = torch.zeros((B, 1, Np, width), dtype=torch.float32, device=dev)
        buf[:, 0, :N, :t.shape[-1]].copy_(t.reshape(B, N, t.shape[-1]))
        return buf

    qp, kp = padded(q.t, d + 1), padded(k.t, d + 1)
    qp[..., d] = 1.0
    kp[:, 0, N:, d] = PAM_KEY_MASK
    sub = Tape(record=tape.record)
    vars_p = [Var(padded(x.t, Cc), False), Var(qp, q.needs_grad), Var(kp, k.needs_grad), Var(padded(v.t, Cc), v.needs_grad)]
    yp = op_pam_core(sub, *vars_p, gamma, precision=precision)
    if out is None:
        out = Var(torch.empty(x.t.shape, dtype=torch.float32, device=dev))
    out.t.view(B, N, Cc).copy_(yp.t[:, 0, :N])          # view(): a channel slice of a concat buffer merges H and W without a copy
    y = out

    def bwd():
        if y.g is None:
            return
        dy = y.g
        yp.g = padded(dy, Cc)
        sub.backward()                       # dq / dk / dv of the padded problem; gamma's gradient lands on the shared Var
        for var, pv, width in ((q, vars_p[1], d), (k, vars_p[2], d), (v, vars_p[3], Cc)):
            if var.needs_grad and pv.g is not None:
                var.add_grad(pv.g[:, 0, :N, :width].reshape(B, H, W, width).contiguous())
        if x.needs_grad:
            _accumulate(x, dy)

    tape.push(bwd)
    return y


def _accumulate(x: Var, g: Tensor) -> None:
    """x.g += g without taking ownership of g (g stays a gradient buffer of someone else)."""
    tgt, acc = x.grad_target()
    axpy(g, tgt, 1.0, acc)


cam_tensor_core: Optional[bool] = None     # None: follow conv_precision (tensor cores unless 'fp32'); True / False force it


def _cam_tc(Cc: int) -> bool:
    use = cam_tensor_core if cam_tensor_core is not None else conv_precision in _PREC
    return bool(use) and Cc % 4 == 0


def op_cam(tape: Tape, x: Var, gamma: Var, *, out: Optional[Var] = None, y16: Optional[Tensor] = None) -> Var:
    """Channel attention: y = gamma * softmax(rowmax(E)-E) X + x with E = X X^T  (generator.py:128-139)."""
    lib = _lib(x.t)
    B, H, W, Cc = x.t.shape
    N = H * W
    dev = x.t.device
    if out is None:
        out = Var(torch.empty(x.t.shape, dtype=torch.float32, device=dev))
    attn = torch.empty((B, Cc, Cc), dtype=torch.float32, device=dev)
    tc = _cam_tc(Cc)
    if y16 is not None:     # output only as a bf16 column block of the consumer's packed operand (see op_pam_core)
        assert tc and y16.dtype == torch.bfloat16 and y16.shape == (B * N, Cc) and y16.stride(1) == 1
        buf = workspace("cam_tc", lib.gdn_cam_tc_ws_bytes(B, N, Cc), dev)
        L.check(lib.gdn_cam_fwd_tc16(x.t.data_ptr(), pitch_of(x.t), gamma.t.data_ptr(), attn.data_ptr(), None, 0, y16.data_ptr(), y16.stride(0), B, N, Cc,
                                     buf.data_ptr(), buf.numel(), _stream()), "gdn_cam_fwd_tc16")
    elif tc:
        buf = workspace("cam_tc", lib.gdn_cam_tc_ws_bytes(B, N, Cc), dev)
        L.check(lib.gdn_cam_fwd_tc(x.t.data_ptr(), pitch_of(x.t), gamma.t.data_ptr(), attn.data_ptr(), out.t.data_ptr(), pitch_of(out.t), B, N, Cc,
                                   buf.data_ptr(), buf.numel(), _stream()), "gdn_cam_fwd_tc")
    else:
        L.check(lib.gdn_cam_fwd(x.t.data_ptr(), pitch_of(x.t), gamma.t.data_ptr(), attn.data_ptr(), out.t.data_ptr(), pitch_of(out.t), B, N, Cc, _stream()), "gdn_cam_fwd")
    y = out

    def bwd():
        if y.g is None:
            return
        dy = y.g
        dgamma = torch.empty(1, dtype=torch.float32, device=dev)
        tgt, acc = x.grad_target()
        if tc:
            buf = workspace("cam_tc", lib.gdn_cam_tc_ws_bytes(B, N, Cc), dev)
            L.check(lib.gdn_cam_bwd_tc(x.t.data_ptr(), pitch_of(x.t), gamma.t.data_ptr(), attn.data_ptr(), dy.data_ptr(), pitch_of(dy), tgt.data_ptr(), pitch_of(tgt),
                                       int(acc), dgamma.data_ptr(), B, N, Cc, buf.data_ptr(), buf.numel(), dot_ws(dev).data_ptr(), _stream()), "gdn_cam_bwd_tc")
        else:
            need = lib.gdn_cam_bwd_ws_bytes(B, N, Cc)
            buf = workspace("cam", need, dev)
            L.check(lib.gdn_cam_bwd(x.t.data_ptr(), pitch_of(x.t), gamma.t.data_ptr(), attn.data_ptr(), dy.data_ptr(), pitch_of(dy), tgt.data_ptr(), pitch_of(tgt),
                                    int(acc), dgamma.data_ptr(), B, N, Cc, buf.data_ptr(), buf.numel(), dot_ws(dev).data_ptr(), _stream()), "gdn_cam_bwd")
        if gamma.needs_grad:
            gamma.add_grad(dgamma)

    tape.push(bwd)
    return y


upsample_skip_fusion: bool = os.environ.get("GDN_UP2_SKIP_FUSION", "1") != "0"


def op_bicubic_up2(tape: Tape, x: Var, skip: Optional[Var] = None) -> Var:
    """nn.Upsample(scale_factor=2, mode='bicubic') (generator.py:221,225).  ``skip``: y = up2(x) + bilinear_resize(skip -> y's grid) in the same pass
    (the skip fusion of generator.py:242-246 on the last up-sampling: the full-resolution tensor is written once)."""
    lib = _lib(x.t)
    B, H, W, Cc = x.t.shape
    assert x.t.is_contiguous()
    y = Var(new_nhwc(B, 2 * H, 2 * W, Cc, x.t))
    if skip is not None:
        _, Hs, Ws, Cs = skip.t.shape
        assert Cs == Cc and skip.t.is_contiguous() and Cc % 4 == 0
        L.check(lib.gdn_bicubic_up2_bilinear_add_fwd(x.t.data_ptr(), skip.t.data_ptr(), y.t.data_ptr(), B, H, W, Hs, Ws, Cc, _stream()), "gdn_bicubic_up2_bilinear_add_fwd")
    else:
        L.check(lib.gdn_bicubic_up2_fwd(x.t.data_ptr(), y.t.data_ptr(), B, H, W, Cc, _stream()), "gdn_bicubic_up2_fwd")

    def bwd():
        if y.g is None:
            return
        assert y.g.is_contiguous()
        if skip is not None and skip.needs_grad:
            tgt, acc = skip.grad_target()
            assert tgt.is_contiguous()
            L.check(lib.gdn_bilinear_bwd(y.g.data_ptr(), tgt.data_ptr(), B, Hs, Ws, 2 * H, 2 * W, Cc, int(acc), _stream()), "gdn_bilinear_bwd")
        if not x.needs_grad:
            return
        gx = new_nhwc(B, H, W, Cc, x.t)
        L.check(lib.gdn_bicubic_up2_bwd(y.g.data_ptr(), gx.data_ptr(), B, H, W, Cc, _stream()), "gdn_bicubic_up2_bwd")
        x.add_grad(gx)

    tape.push(bwd)
    return y


head_tap_planes: bool = os.environ.get("GDN_HEAD_TAP_PLANES", "1") != "0"
HEAD_T = 12      # tap planes per pixel: 9 used, padded to a multiple of 4 for the 128-bit resampling kernels


def head_tap_planes_ok(u: Var, s: Optional[Var], w: Var) -> bool:
    O, Cc, kh, kw = w.t.shape
    return bool(head_tap_planes and O == 1 and kh == 3 and kw == 3 and u.t.shape[-1] == Cc and u.t.is_contiguous()
                and (s is None or (s.t.shape[-1] == Cc and s.t.is_contiguous())))


def op_upsample_skip_final(tape: Tape, u: Var, s: Optional[Var], w: Var, bias: Optional[Var]) -> Var:
    """``final(up2(u) + bilinear_resize(s))`` (generator.py:225,242-246,228: last bicubic up-sampling, skip fusion, 3x3 convolution C -> 1, padding 1)
    WITHOUT the C-channel full-resolution tensor.  Everything here is linear and the resamplers act per channel, so the convolution's channel
    reduction is hoisted in front of them: z_t = sum_c w[c][t] u_c (a 1x1 convolution to 9 tap planes at u's resolution; likewise for s),
    Z = up2(z) + resize(zs), y[p] = bias + sum_t Z_t[p + t] with zero padding.  Exact in real arithmetic (another fp32 summation order);
    the full-resolution traffic drops from 64 to 12 channels in the forward and the backward (1.07 GB -> 0.2 GB per pass at B = 32, 256x512)."""
    lib = _lib(u.t)
    B, H, W, Cc = u.t.shape
    dev = u.t.device
    T = HEAD_T
    wp = torch.zeros((T, Cc), dtype=torch.float32, device=dev)
    wp[:9].copy_(w.t.detach().reshape(Cc, 9).t())                      # wp[kh*3+kw][c] = w[0][c][kh][kw]
    wpt = wp.t().contiguous()                                           # [C, T]: the adjoint 1x1 convolution's weight
    narrow = Cc % 4 == 0 and Cc <= 256 and 256 % (Cc // 4) == 0        # HBM-bound narrow 1x1 kernels (thin_conv.cu); else the fp32 implicit-GEMM engine

    def c1x1(x_t: Tensor, out: Tensor) -> None:
        if narrow:
            L.check(lib.gdn_narrow_conv1x1_fwd(x_t.data_ptr(), Cc, wp.data_ptr(), T, out.data_ptr(), rows_of(x_t), _stream()), "gdn_narrow_conv1x1_fwd")
        else:
            conv_raw(x_t, wp.view(T, 1, 1, Cc), out, kh=1, kw=1)

    def c1x1_dgrad(dz_t: Tensor, tgt: Tensor, acc: bool) -> None:
        if narrow and tgt.is_contiguous():
            L.check(lib.gdn_narrow_conv1x1_dgrad(dz_t.data_ptr(), T, wp.data_ptr(), Cc, tgt.data_ptr(), rows_of(dz_t), int(acc), _stream()), "gdn_narrow_conv1x1_dgrad")
        else:
            conv_raw(dz_t, wpt.view(Cc, 1, 1, T), tgt, kh=1, kw=1, res=tgt if acc else None)

    def c1x1_wgrad(dz_t: Tensor, x_t: Tensor, out: Tensor, acc: bool) -> None:
        if narrow:
            M = rows_of(x_t)
            buf = workspace("narrow_wgrad", lib.gdn_narrow_conv1x1_wgrad_ws_bytes(M, Cc, T), dev)
            L.check(lib.gdn_narrow_conv1x1_wgrad(dz_t.data_ptr(), T, x_t.data_ptr(), Cc, M, out.data_ptr(), int(acc), buf.data_ptr(), buf.numel(), _stream()),
                    "gdn_narrow_conv1x1_wgrad")
        else:
            wgrad_raw(dz_t, x_t, out, kh=1, kw=1, accumulate=acc)

    z = new_nhwc(B, H, W, T, u.t)
    c1x1(u.t, z)
    Z = new_nhwc(B, 2 * H, 2 * W, T, u.t)
    if s is not None:
        _, Hs, Ws, _ = s.t.shape
        zs = new_nhwc(B, Hs, Ws, T, u.t)
        c1x1(s.t, zs)
        L.check(lib.gdn_bicubic_up2_bilinear_add_fwd(z.data_ptr(), zs.data_ptr(), Z.data_ptr(), B, H, W, Hs, Ws, T, _stream()), "gdn_bicubic_up2_bilinear_add_fwd")
    else:
        L.check(lib.gdn_bicubic_up2_fwd(z.data_ptr(), Z.data_ptr(), B, H, W, T, _stream()), "gdn_bicubic_up2_fwd")
    y = Var(new_nhwc(B, 2 * H, 2 * W, 1, u.t))
    L.check(lib.gdn_tap_shift_sum(Z.data_ptr(), T, None if bias is None else bias.t.data_ptr(), y.t.data_ptr(), B, 2 * H, 2 * W, _stream()), "gdn_tap_shift_sum")
    del z, Z

    def bwd():
        if y.g is None:
            return
        dy = y.g
        assert dy.is_contiguous()
        dZ = new_nhwc(B, 2 * H, 2 * W, T, u.t)
        L.check(lib.gdn_tap_shift_expand(dy.data_ptr(), dZ.data_ptr(), T, B, 2 * H, 2 * W, _stream()), "gdn_tap_shift_expand")
        dz = new_nhwc(B, H, W, T, u.t)
        L.check(lib.gdn_bicubic_up2_bwd(dZ.data_ptr(), dz.data_ptr(), B, H, W, T, _stream()), "gdn_bicubic_up2_bwd")
        gwp = None
        if w.needs_grad:
            gwp = torch.empty((T, Cc, 1, 1), dtype=torch.float32, device=dev)
            c1x1_wgrad(dz, u.t, gwp, False)
        if u.needs_grad:
            tgt, acc = u.grad_target()
            c1x1_dgrad(dz, tgt, acc)
        if s is not None and (s.needs_grad or w.needs_grad):
            dzs = new_nhwc(B, Hs, Ws, T, u.t)
            L.check(lib.gdn_bilinear_bwd(dZ.data_ptr(), dzs.data_ptr(), B, Hs, Ws, 2 * H, 2 * W, T, 0, _stream()), "gdn_bilinear_bwd")
            if w.needs_grad:
                c1x1_wgrad(dzs, s.t, gwp, True)
            if s.needs_grad:
                tgt, acc = s.grad_target()
                c1x1_dgrad(dzs, tgt, acc)
        if gwp is not None:
            w.add_grad(gwp.view(T, Cc)[:9].t().reshape(1, Cc, 3, 3).contiguous())
        if bias is not None and bias.needs_grad:
            bias.add_grad(colsum_f32(dy.view(-1, 1)))

    tape.push(bwd)
    return y


def op_bilinear_add_(tape: Tape, s: Var, x: Var) -> Var:
    """x += bilinear_resize(s, x.shape) in place; returns the Var standing for the sum (shares x's storage)."""
    lib = _lib(x.t)
    B, Hi, Wi, Cc = s.t.shape
    _, Ho, Wo, _ = x.t.shape
    assert s.t.is_contiguous() and x.t.is_contiguous()
    L.check(lib.gdn_bilinear_fwd(s.t.data_ptr(), x.t.data_ptr(), B, Hi, Wi, Ho, Wo, Cc, 1, _stream()), "gdn_bilinear_fwd")
    y = Var(x.t)

    def bwd():
        if y.g is None:
            return
        if s.needs_grad:
            tgt, acc = s.grad_target()
            assert tgt.is_contiguous()
            L.check(lib.gdn_bilinear_bwd(y.g.data_ptr(), tgt.data_ptr(), B, Hi, Wi, Ho, Wo, Cc, int(acc), _stream()), "gdn_bilinear_bwd")
        if x.needs_grad:
            x.add_grad(y.g)

    tape.push(bwd)
    return y


def op_conv_accumulate(tape: Tape, x: Var, w: Var, acc: Optional[Var]) -> Var:
    """acc (+)= conv1x1(x, w) (no bias); returns acc (allocated when None).  Used for the hoisted skip projections."""
    O, I, kh, kw = w.t.shape
    B, H, W, Cin = x.t.shape
    first = acc is None
    if first:
        acc = Var(new_nhwc(B, H, W, O, x.t))
    cctx = conv_forward(x.t, w.t.detach(), acc.t, res=None if first else acc.t, keep=tape.record and w.needs_grad)
    y = acc

    def bwd():
        if y.g is None:
            return
        gw = torch.empty_like(w.t) if w.needs_grad else None
        tgt, a2 = x.grad_target() if x.needs_grad else (None, False)
        conv_backward(cctx, y.g, x.t, w.t.detach(), gw=gw, gx=tgt, gx_accumulate=a2)
        if gw is not None:
            w.add_grad(gw)

    tape.push(bwd)
    return y


def op_maxpool2(tape: Tape, x: Var) -> Var:
    lib = _lib(x.t)
    B, H, W, Cc = x.t.shape
    assert x.t.is_contiguous()
    y = Var(new_nhwc(B, H // 2, W // 2, Cc, x.t))
    L.check(lib.gdn_maxpool2_fwd(x.t.data_ptr(), y.t.data_ptr(), B, H, W, Cc, _stream()), "gdn_maxpool2_fwd")

    def bwd():
        if y.g is None or not x.needs_grad:
            return
        gx = new_nhwc(B, H, W, Cc, x.t)
        L.check(lib.gdn_maxpool2_bwd(x.t.data_ptr(), y.g.data_ptr(), gx.data_ptr(), B, H, W, Cc, _stream()), "gdn_maxpool2_bwd")
        x.add_grad(gx)

    tape.push(bwd)
    return y


pool_relu_pack_fusion: bool = os.environ.get("GDN_POOL_RELU_PACK", "1") != "0"


def op_maxpool2_relu_packed(tape: Tape, x: Var) -> Var:
    """MaxPool2d(2, 2) of a conv + ReLU output that feeds ONLY this pool and whose convolution takes its data gradient as a packed bf16 operand: the
    backward writes x.g16 = bf16(route(dy) * (x > 0)) in one pass (gdn_maxpool2_bwd_relu_pack16) instead of an fp32 dx followed by pack_actgrad."""
    lib = _lib(x.t)
    B, H, W, Cc = x.t.shape
    if not (pool_relu_pack_fusion and conv_precision == "bf16" and H % 2 == 0 and W % 2 == 0 and Cc % 8 == 0 and x.t.is_contiguous()):
        return op_maxpool2(tape, x)
    y = Var(new_nhwc(B, H // 2, W // 2, Cc, x.t))
    L.check(lib.gdn_maxpool2_fwd(x.t.data_ptr(), y.t.data_ptr(), B, H, W, Cc, _stream()), "gdn_maxpool2_fwd")

    def bwd():
        if y.g is None or not x.needs_grad:
            return
        if x.g is not None or x.parent is not None or x.g16 is not None:      # another consumer already produced a gradient: the generic route
            gx = new_nhwc(B, H, W, Cc, x.t)
            L.check(lib.gdn_maxpool2_bwd(x.t.data_ptr(), y.g.data_ptr(), gx.data_ptr(), B, H, W, Cc, _stream()), "gdn_maxpool2_bwd")
            x.add_grad(gx)
            return
        dz16 = torch.empty((B * H * W, Cc), dtype=torch.bfloat16, device=x.t.device)
        L.check(lib.gdn_maxpool2_bwd_relu_pack16(x.t.data_ptr(), y.g.data_ptr(), dz16.data_ptr(), B, H, W, Cc, _stream()), "gdn_maxpool2_bwd_relu_pack16")
        x.g16 = Packed(dz16, None, Cc)

    tape.push(bwd)
    return y


def op_maxpool2_bf16(tape: Tape, x: Var, x16: Tensor) -> Tuple[Var, Tensor]:
    """MaxPool2d(2, 2) on a bf16-only feature map (x.t is a shape-only placeholder; the gradients of x and y stay fp32)."""
    lib = _lib(x16)
    B, H, W, Cc = x.t.shape
    y = Var(new_nhwc(B, H // 2, W // 2, Cc, x.t), needs_grad=x.needs_grad)
    y16 = torch.empty((B * (H // 2) * (W // 2), Cc), dtype=torch.bfloat16, device=x16.device)
    L.check(lib.gdn_maxpool2_fwd_bf16(x16.data_ptr(), y16.data_ptr(), B, H, W, Cc, _stream()), "gdn_maxpool2_fwd_bf16")

    def bwd():
        if y.g is None or not x.needs_grad:
            return
        gx = new_nhwc(B, H, W, Cc, x.t)
        L.check(lib.gdn_maxpool2_bwd_bf16(x16.data_ptr(), y.g.data_ptr(), gx.data_ptr(), B, H, W, Cc, _stream()), "gdn_maxpool2_bwd_bf16")
        x.add_grad(gx)

    tape.push(bwd)
    return y, y16


def op_cat_rows(tape: Tape, parts: List[Var]) -> Var:
    """Concatenation of parameter tensors along dim 0 (tiny: the weights / biases of projections that run as ONE convolution);
    the backward hands every part its rows of the gradient."""
    y = Var(torch.cat([p.t.detach() for p in parts], dim=0), any(p.needs_grad for p in parts))

    def bwd():
        if y.g is None:
            return
        r = 0
        for p in parts:
            n = p.t.shape[0]
            if p.needs_grad:
                p.add_grad(y.g[r:r + n].contiguous())
            r += n

    tape.push(bwd)
    return y


pam_merge_qk: bool = os.environ.get("GDN_PAM_MERGE_QK", "1") != "0"


def op_copy(tape: Tape, x: Var, out: Var) -> Var:
    """out = x (into a slice of a wider buffer)."""
    axpy(x.t, out.t, 1.0, False)

    def bwd():
        if out.g is not None and x.needs_grad:
            _accumulate(x, out.g)

    tape.push(bwd)
    return out


def op_from_nchw(tape: Tape, x_nchw: Tensor, needs_grad: bool, out: Optional[Var] = None) -> Var:
    B, Cc, H, W = x_nchw.shape
    if out is None:
        out = Var(new_nhwc(B, H, W, Cc, x_nchw), needs_grad)
    else:
        out.needs_grad = needs_grad
    _f32(x_nchw)
    if Cc == 1:
        axpy(x_nchw.contiguous().view(B, H, W, 1), out.t, 1.0, False)
    else:
        nchw_to_nhwc(x_nchw.contiguous(), out.t)
    return out


def to_nchw(x: Tensor) -> Tensor:
    B, H, W, Cc = x.shape
    out = torch.empty((B, Cc, H, W), dtype=torch.float32, device=x.device)
    if Cc == 1 and x.is_contiguous():
        axpy(x, out.view(B, H, W, 1), 1.0, False)
    else:
        nhwc_to_nchw(x, out)
    return out


def grad_to_nhwc(g_nchw: Tensor) -> Tensor:
    B, Cc, H, W = g_nchw.shape
    out = torch.empty((B, H, W, Cc), dtype=torch.float32, device=g_nchw.device)
    g = g_nchw.contiguous()
    if Cc == 1:
        axpy(g.view(B, H, W, 1), out, 1.0, False)
    else:
        nchw_to_nhwc(g, out)
    return out


# ----------------------------------------------------------------------------------------------------------------
# bridge to torch.autograd
# ----------------------------------------------------------------------------------------------------------------


class TapeFunction(torch.autograd.Function):
    """Runs ``build(tape, x_nchw, x_needs_grad, param_vars) -> (out Var (NHWC or [B,F]), input Var)`` and exposes it
    as one autograd node.  apply(build, x, *params): ``params`` are the leaf tensors whose gradients the tape produces."""

    @staticmethod
    def forward(ctx, build, x: Tensor, *params: Tensor):
        needs = ctx.needs_input_grad            # (build, x, *params)
        record = any(needs)
        tape = Tape(record)
        pvars = [Var(p.detach(), bool(needs[2 + i])) for i, p in enumerate(params)]
        if not x.is_cuda:
            raise L.GdnError("gan_danet_b200 modules need CUDA tensors on an sm_100 device (no CPU fallback)")
        with torch.cuda.device(x.device):
            out, xin = build(tape, x.detach(), bool(needs[1]), pvars)
        ctx.tape, ctx.pvars, ctx.xin, ctx.out = tape, pvars, xin, out
        ctx.x_shape, ctx.dev = x.shape, x.device
        ctx.params, ctx.versions = params, [p._version for p in params]      # an in-place update between forward and backward would be read by dgrad
        res = to_nchw(out.t) if out.t.dim() == 4 else out.t
        return res

    @staticmethod
    def backward(ctx, dout: Tensor):
        if ctx.tape is None or ctx.tape.consumed:
            raise L.GdnError("backward through a gan_danet_b200 module a second time: its saved buffers are released after the first backward "
                             "(retain_graph=True / two separate backward passes through one forward are not supported -- sum the losses first)")
        changed = [i for i, (p, v) in enumerate(zip(ctx.params, ctx.versions)) if p._version != v]
        if changed:
            raise L.GdnError(f"{len(changed)} parameter tensor(s) were modified in place between this module's forward and backward "
                             "(e.g. optimizer.step() before .backward()): the data gradient would use the updated weights")
        out = ctx.out
        with torch.cuda.device(ctx.dev):
            dout = dout.contiguous()
            out.g = grad_to_nhwc(dout) if out.t.dim() == 4 else dout
            ctx.tape.backward()
            gx = None
            if ctx.needs_input_grad[1] and ctx.xin.g is not None:
                gx = to_nchw(ctx.xin.g) if ctx.xin.g.dim() == 4 else ctx.xin.g
            grads = []
            for i, pv in enumerate(ctx.pvars):
                if ctx.needs_input_grad[2 + i]:
                    g = pv.g
                    if g is None:
                        g = torch.empty_like(pv.t)
                        fill_(g.view(-1) if g.is_contiguous() else g, 0.0)
                    grads.append(g.view(pv.t.shape))
                else:
                    grads.append(None)
        ctx.pvars = ctx.xin = ctx.out = ctx.params = None       # ctx.tape stays (consumed) so that a second backward raises instead of returning zeros
        return (None, gx, *grads)


def op_add_(tape: Tape, y: Var, r: Var) -> Var:
    """y += r in place (residual add); gradients flow to both."""
    axpy(r.t, y.t, 1.0, True)

    def bwd():
        if y.g is not None and r.needs_grad:
            _accumulate(r, y.g)

    tape.push(bwd)
    return y


def op_global_avg_pool(tape: Tape, x: Var) -> Var:
    """nn.AdaptiveAvgPool2d((1,1)) + flatten: [B,H,W,C] -> [B,C]."""
    B, H, W, Cc = x.t.shape
    y = Var(torch.empty((B, Cc), dtype=torch.float32, device=x.t.device))
    inv = 1.0 / float(H * W)
    for b in range(B):
        s = colstats(x.t[b])
        L.check(_lib(x.t).gdn_sums_to_float(s.data_ptr(), y.t[b].data_ptr(), Cc, inv, _stream()), "gdn_sums_to_float")

    def bwd():
        if y.g is None or not x.needs_grad:
            return
        gx = new_nhwc(B, H, W, Cc, x.t)
        zero = torch.empty(Cc, dtype=torch.float32, device=x.t.device)
        fill_(zero, 0.0)
        shift = torch.empty((B, Cc), dtype=torch.float32, device=x.t.device)
        axpy(y.g, shift, inv, False)
        for b in range(B):
            affine_act(x.t[b], gx[b], zero, shift[b], ACT_NONE)
        x.add_grad(gx)

    tape.push(bwd)
    return y
