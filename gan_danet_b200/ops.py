"""``torch.ops.gandanet.*`` -- the custom-op layer over the C ABI (SURVEY 8b "custom-op layer"; north_star: "PyTorch custom ops
(torch.library plus a thin C++/C-ABI extension)").

Each op is registered with ``torch.library.custom_op`` for the **CUDA dispatch key only** (a CPU tensor raises
``NotImplementedError``: there is no CPU kernel, by design), has a fake (meta) implementation for shape inference /
``torch.compile`` tracing, and -- where the reference differentiates through it -- an autograd formula that calls the matching
``*_bwd`` op.  The implementations marshal raw pointers into ``libgandanet_sm100.so`` through ``ctypes`` on PyTorch's current
stream; nothing here computes on the host.  Layout: activations are NHWC float32 (``x.permute(0, 2, 3, 1)`` of the reference's
NCHW tensors); ``gamma`` is the 1-element parameter of the attention modules.

    y, o, lse = torch.ops.gandanet.pam_fwd(x, q, k, v, gamma, "fp16")      # generator.py:115-122, fused tcgen05 flash kernel
    y, attn   = torch.ops.gandanet.cam_fwd(x, gamma, True)                 # generator.py:131-139
    y         = torch.ops.gandanet.conv2d(x, w, bias, 1, 1, 1, 0.0)        # nn.Conv2d (+ ReLU), OIHW weight
    y         = torch.ops.gandanet.upsample_bicubic2x(x)                   # nn.Upsample(2, 'bicubic'), generator.py:221,225
    torch.ops.gandanet.fused_adamw_(p, g, m, v, lr, b1, b2, eps, wd, step, 1.0)   # optim.AdamW.step, GAN_DANet_train.ipynb:182-183

The ``nn.Module`` mirror in ``gan_danet_b200.models`` records the same kernels on a tape (one autograd node per module); this
layer serves callers that keep the reference's own modules and swap single ops (level 2).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib as L
from . import engine as E

_PRECISIONS = {"fp32": L.PREC_FP32, "fp16": L.PREC_FP16, "fp16x3": L.PREC_FP16X3}


def _nhwc(t: Tensor, what: str) -> Tensor:
    if t.dim() != 4 or t.dtype != torch.float32:
        raise L.GdnError(f"{what}: expected a 4-d float32 NHWC tensor, got {tuple(t.shape)} {t.dtype}")
    return t.contiguous()


def _pam_precision(precision: str, N: int, d: int, Cc: int) -> int:
    if precision not in _PRECISIONS:
        raise L.GdnError(f"pam precision {precision!r}: expected 'fp16x3' / 'fp16' (fused tcgen05 kernels, split / single fp16 logit operands) or 'fp32' (parity engine)")
    p = _PRECISIONS[precision]
    if p != L.PREC_FP32 and (N % 128 != 0 or d > 30 or Cc >= 191 or Cc % 4 != 0):
        # No silent fall-back: the fp32 CUDA-core engine is ~15x slower, a cliff the caller must choose.  (The module path pads grids with
        # N % 128 != 0 per sample and masks the padded keys: engine._op_pam_core_padded.)
        raise L.GdnError(f"gandanet::pam_fwd/pam_bwd precision {precision!r}: shape N={N} d={d} C={Cc} is outside the fused tensor-core kernels "
                         "(N % 128 == 0, d <= 30, C <= 188, C % 4 == 0); pad the grid to a multiple of 128 positions (PAMModule does) or pass precision='fp32'")
    return p


# ------------------------------------------------------------------------------------------------ position attention
@torch.library.custom_op("gandanet::pam_fwd", mutates_args=(), device_types="cuda")
def pam_fwd(x: Tensor, q: Tensor, k: Tensor, v: Tensor, gamma: Tensor, precision: str) -> Tuple[Tensor, Tensor, Tensor]:
    """y = gamma * softmax(q k^T) v + x over the N = H*W positions (generator.py:115-122); also returns the normalised
    attention output o [B, N, C] and the row log-sum-exp lse [B, N] that the backward needs (the N x N map is never stored)."""
    x, q, k, v = _nhwc(x, "pam_fwd(x)"), _nhwc(q, "pam_fwd(q)"), _nhwc(k, "pam_fwd(k)"), _nhwc(v, "pam_fwd(v)")
    lib = E._lib(x)
    B, H, W, Cc = x.shape
    N, d = H * W, q.shape[-1]
    y = torch.empty_like(x)
    o = torch.empty((B, N, Cc), dtype=torch.float32, device=x.device)
    lse = torch.empty((B, N), dtype=torch.float32, device=x.device)
    a = L.PamFwdArgs()
    a.q, a.k, a.qk_pitch, a.d = q.data_ptr(), k.data_ptr(), d, d
    a.v, a.v_pitch = v.data_ptr(), Cc
    a.x, a.x_pitch, a.gamma = x.data_ptr(), Cc, gamma.data_ptr()
    a.o, a.y, a.y_pitch, a.lse = o.data_ptr(), y.data_ptr(), Cc, lse.data_ptr()
    a.B, a.N, a.C, a.precision, a.chunk = B, N, Cc, _pam_precision(precision, N, d, Cc), 0
    ws = E.workspace("pam", lib.gdn_pam_fwd_ws_bytes(C.byref(a)), x.device)
    a.ws, a.ws_bytes = ws.data_ptr(), ws.numel()
    L.check(lib.gdn_pam_fwd(C.byref(a), E._stream()), "gdn_pam_fwd")
    return y, o, lse


@pam_fwd.register_fake
def _(x, q, k, v, gamma, precision):
    B, H, W, Cc = x.shape
    return torch.empty_like(x), x.new_empty((B, H * W, Cc)), x.new_empty((B, H * W))


@torch.library.custom_op("gandanet::pam_bwd", mutates_args=(), device_types="cuda")
def pam_bwd(dy: Tensor, q: Tensor, k: Tensor, v: Tensor, o: Tensor, lse: Tensor, gamma: Tensor, precision: str) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """(dq, dk, dv, dgamma) of pam_fwd for the cotangent dy of y; dx = dy is the caller's identity term."""
    dy, q, k, v = _nhwc(dy, "pam_bwd(dy)"), _nhwc(q, "pam_bwd(q)"), _nhwc(k, "pam_bwd(k)"), _nhwc(v, "pam_bwd(v)")
    lib = E._lib(dy)
    B, H, W, Cc = dy.shape
    N, d = H * W, q.shape[-1]
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    rowdot = torch.empty((B * N, 1), dtype=torch.float32, device=dy.device)
    b = L.PamBwdArgs()
    b.q, b.k, b.qk_pitch, b.d = q.data_ptr(), k.data_ptr(), d, d
    b.v, b.v_pitch = v.data_ptr(), Cc
    b.o, b.lse, b.gamma = o.contiguous().data_ptr(), lse.contiguous().data_ptr(), gamma.data_ptr()
    b.dy, b.dy_pitch = dy.data_ptr(), Cc
    b.dq, b.dk, b.dv, b.rowdot = dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), rowdot.data_ptr()
    prec = _pam_precision(precision, N, d, Cc)
    b.B, b.N, b.C, b.precision, b.chunk = B, N, Cc, (prec if E.pam_bwd_tensor_core else L.PREC_FP32), 0
    ws = E.workspace("pam", lib.gdn_pam_bwd_ws_bytes(C.byref(b)), dy.device)
    b.ws, b.ws_bytes = ws.data_ptr(), ws.numel()
    L.check(lib.gdn_pam_bwd(C.byref(b), E._stream()), "gdn_pam_bwd")
    dgamma = E.sums_to_float(E.colstats(rowdot), 1)
    return dq, dk, dv, dgamma


@pam_bwd.register_fake
def _(dy, q, k, v, o, lse, gamma, precision):
    return torch.empty_like(q), torch.empty_like(k), torch.empty_like(v), gamma.new_empty((1,))


def _pam_setup(ctx, inputs, output):
    x, q, k, v, gamma, precision = inputs
    _, o, lse = output
    ctx.save_for_backward(q, k, v, o, lse, gamma)
    ctx.precision = precision


def _pam_backward(ctx, dy, do, dlse):
    q, k, v, o, lse, gamma = ctx.saved_tensors
    dq, dk, dv, dgamma = torch.ops.gandanet.pam_bwd(dy.contiguous(), q, k, v, o, lse, gamma, ctx.precision)
    return dy, dq, dk, dv, dgamma.reshape(gamma.shape), None      # o and lse are saved statistics, not differentiable outputs


pam_fwd.register_autograd(_pam_backward, setup_context=_pam_setup)


# ------------------------------------------------------------------------------------------------ channel attention
@torch.library.custom_op("gandanet::cam_fwd", mutates_args=(), device_types="cuda")
def cam_fwd(x: Tensor, gamma: Tensor, tensor_core: bool) -> Tuple[Tensor, Tensor]:
    """y = gamma * softmax(rowmax(E) - E) X + x with E = X X^T over the positions (generator.py:128-139); attn [B, C, C] is returned
    for the backward.  ``tensor_core``: Gram matrix and re-projection on tcgen05 with bf16 hi+lo split operands (needs C % 4 == 0)."""
    x = _nhwc(x, "cam_fwd(x)")
    lib = E._lib(x)
    B, H, W, Cc = x.shape
    N = H * W
    y = torch.empty_like(x)
    attn = torch.empty((B, Cc, Cc), dtype=torch.float32, device=x.device)
    if tensor_core and Cc % 4 == 0:
        ws = E.workspace("cam_tc", lib.gdn_cam_tc_ws_bytes(B, N, Cc), x.device)
        L.check(lib.gdn_cam_fwd_tc(x.data_ptr(), Cc, gamma.data_ptr(), attn.data_ptr(), y.data_ptr(), Cc, B, N, Cc, ws.data_ptr(), ws.numel(), E._stream()), "gdn_cam_fwd_tc")
    else:
        L.check(lib.gdn_cam_fwd(x.data_ptr(), Cc, gamma.data_ptr(), attn.data_ptr(), y.data_ptr(), Cc, B, N, Cc, E._stream()), "gdn_cam_fwd")
    return y, attn


@cam_fwd.register_fake
def _(x, gamma, tensor_core):
    return torch.empty_like(x), x.new_empty((x.shape[0], x.shape[3], x.shape[3]))


@torch.library.custom_op("gandanet::cam_bwd", mutates_args=(), device_types="cuda")
def cam_bwd(dy: Tensor, x: Tensor, gamma: Tensor, attn: Tensor, tensor_core: bool) -> Tuple[Tensor, Tensor]:
    """(dx, dgamma) of cam_fwd, including the identity term of the residual."""
    dy, x = _nhwc(dy, "cam_bwd(dy)"), _nhwc(x, "cam_bwd(x)")
    lib = E._lib(x)
    B, H, W, Cc = x.shape
    N = H * W
    dx = torch.empty_like(x)
    dgamma = torch.empty(1, dtype=torch.float32, device=x.device)
    attn = attn.contiguous()
    dot = E.dot_ws(x.device)
    if tensor_core and Cc % 4 == 0:
        ws = E.workspace("cam_tc", lib.gdn_cam_tc_ws_bytes(B, N, Cc), x.device)
        L.check(lib.gdn_cam_bwd_tc(x.data_ptr(), Cc, gamma.data_ptr(), attn.data_ptr(), dy.data_ptr(), Cc, dx.data_ptr(), Cc, 0, dgamma.data_ptr(), B, N, Cc,
                                   ws.data_ptr(), ws.numel(), dot.data_ptr(), E._stream()), "gdn_cam_bwd_tc")
    else:
        ws = E.workspace("cam", lib.gdn_cam_bwd_ws_bytes(B, N, Cc), x.device)
        L.check(lib.gdn_cam_bwd(x.data_ptr(), Cc, gamma.data_ptr(), attn.data_ptr(), dy.data_ptr(), Cc, dx.data_ptr(), Cc, 0, dgamma.data_ptr(), B, N, Cc,
                                ws.data_ptr(), ws.numel(), dot.data_ptr(), E._stream()), "gdn_cam_bwd")
    return dx, dgamma


@cam_bwd.register_fake
def _(dy, x, gamma, attn, tensor_core):
    return torch.empty_like(x), gamma.new_empty((1,))


def _cam_setup(ctx, inputs, output):
    x, gamma, tensor_core = inputs
    ctx.save_for_backward(x, gamma, output[1])
    ctx.tensor_core = tensor_core


def _cam_backward(ctx, dy, dattn):
    x, gamma, attn = ctx.saved_tensors
    dx, dgamma = torch.ops.gandanet.cam_bwd(dy.contiguous(), x, gamma, attn, ctx.tensor_core)
    return dx, dgamma.reshape(gamma.shape), None


cam_fwd.register_autograd(_cam_backward, setup_context=_cam_setup)


# ------------------------------------------------------------------------------------------------ convolution
def _conv_out_hw(x: Tensor, w: Tensor, stride: int, pad: int) -> Tuple[int, int]:
    return (x.shape[1] + 2 * pad - w.shape[2]) // stride + 1, (x.shape[2] + 2 * pad - w.shape[3]) // stride + 1


@torch.library.custom_op("gandanet::conv2d", mutates_args=(), device_types="cuda")
def conv2d(x: Tensor, w: Tensor, bias: Optional[Tensor], stride: int, pad: int, act: int, slope: float) -> Tensor:
    """act(conv2d(x, w) + bias) for NHWC x and an OIHW weight (nn.Conv2d of generator.py / discriminator.py / VGG19 with the
    following ReLU / LeakyReLU fused: act 0 none, 1 ReLU, 2 LeakyReLU(slope)).  Runs on the engine selected by
    ``engine.set_conv_precision`` ('bf16' / 'bf16x3': tcgen05 implicit GEMM; 'fp32': CUDA-core parity engine)."""
    x = _nhwc(x, "conv2d(x)")
    Ho, Wo = _conv_out_hw(x, w, stride, pad)
    y = torch.empty((x.shape[0], Ho, Wo, w.shape[0]), dtype=torch.float32, device=x.device)
    E.conv_forward(x, w.contiguous(), y, stride=stride, pad=pad, bias=bias, act=act, slope=slope, keep=False)
    return y


@conv2d.register_fake
def _(x, w, bias, stride, pad, act, slope):
    Ho, Wo = _conv_out_hw(x, w, stride, pad)
    return x.new_empty((x.shape[0], Ho, Wo, w.shape[0]))


@torch.library.custom_op("gandanet::conv2d_bwd", mutates_args=(), device_types="cuda")
def conv2d_bwd(dy: Tensor, x: Tensor, w: Tensor, y: Tensor, stride: int, pad: int, act: int, slope: float, need_dx: bool, need_dw: bool,
               need_db: bool) -> Tuple[Tensor, Tensor, Tensor]:
    """(dx, dw, dbias) of conv2d; a gradient that is not needed comes back as an empty tensor."""
    dy, x, y = _nhwc(dy, "conv2d_bwd(dy)"), _nhwc(x, "conv2d_bwd(x)"), _nhwc(y, "conv2d_bwd(y)")
    w = w.contiguous()
    if act != L.ACT_NONE:
        dz = torch.empty_like(dy)
        E.act_bwd(dy, y, dz, act, slope)
    else:
        dz = dy
    none = torch.empty(0, dtype=torch.float32, device=x.device)
    dx = torch.empty_like(x) if need_dx else None
    dw = torch.empty_like(w) if need_dw else None
    db = E.sums_to_float(E.colstats(dz), w.shape[0]) if need_db else none
    E.conv_backward(E.ConvCtx(False, None), dz, x, w, stride=stride, pad=pad, gw=dw, gx=dx, gx_accumulate=False)
    return (dx if need_dx else none), (dw if need_dw else none), db


@conv2d_bwd.register_fake
def _(dy, x, w, y, stride, pad, act, slope, need_dx, need_dw, need_db):
    return (torch.empty_like(x) if need_dx else x.new_empty(0)), (torch.empty_like(w) if need_dw else x.new_empty(0)), x.new_empty(w.shape[0] if need_db else 0)


def _conv_setup(ctx, inputs, output):
    x, w, bias, stride, pad, act, slope = inputs
    ctx.save_for_backward(x, w, output)
    ctx.cfg = (stride, pad, act, slope, bias is not None)


def _conv_backward(ctx, dy):
    x, w, y = ctx.saved_tensors
    stride, pad, act, slope, has_bias = ctx.cfg
    need = ctx.needs_input_grad
    dx, dw, db = torch.ops.gandanet.conv2d_bwd(dy.contiguous(), x, w, y, stride, pad, act, slope, bool(need[0]), bool(need[1]), bool(has_bias and need[2]))
    return (dx if need[0] else None), (dw if need[1] else None), (db if has_bias and need[2] else None), None, None, None, None


conv2d.register_autograd(_conv_backward, setup_context=_conv_setup)


# ------------------------------------------------------------------------------------------------ bicubic x2
@torch.library.custom_op("gandanet::upsample_bicubic2x", mutates_args=(), device_types="cuda")
def upsample_bicubic2x(x: Tensor) -> Tensor:
    """nn.Upsample(scale_factor=2, mode='bicubic', align_corners=False) on NHWC (generator.py:221,225)."""
    x = _nhwc(x, "upsample_bicubic2x(x)")
    B, H, W, Cc = x.shape
    y = torch.empty((B, 2 * H, 2 * W, Cc), dtype=torch.float32, device=x.device)
    L.check(E._lib(x).gdn_bicubic_up2_fwd(x.data_ptr(), y.data_ptr(), B, H, W, Cc, E._stream()), "gdn_bicubic_up2_fwd")
    return y


@upsample_bicubic2x.register_fake
def _(x):
    B, H, W, Cc = x.shape
    return x.new_empty((B, 2 * H, 2 * W, Cc))


@torch.library.custom_op("gandanet::upsample_bicubic2x_bwd", mutates_args=(), device_types="cuda")
def upsample_bicubic2x_bwd(dy: Tensor) -> Tensor:
    dy = _nhwc(dy, "upsample_bicubic2x_bwd(dy)")
    B, H2, W2, Cc = dy.shape
    dx = torch.empty((B, H2 // 2, W2 // 2, Cc), dtype=torch.float32, device=dy.device)
    L.check(E._lib(dy).gdn_bicubic_up2_bwd(dy.data_ptr(), dx.data_ptr(), B, H2 // 2, W2 // 2, Cc, E._stream()), "gdn_bicubic_up2_bwd")
    return dx


@upsample_bicubic2x_bwd.register_fake
def _(dy):
    B, H2, W2, Cc = dy.shape
    return dy.new_empty((B, H2 // 2, W2 // 2, Cc))


upsample_bicubic2x.register_autograd(lambda ctx, dy: torch.ops.gandanet.upsample_bicubic2x_bwd(dy.contiguous()))


# ------------------------------------------------------------------------------------------------ optimiser
@torch.library.custom_op("gandanet::fused_adamw_", mutates_args=("p", "m", "v"), device_types="cuda")
def fused_adamw_(p: Tensor, g: Tensor, m: Tensor, v: Tensor, lr: float, beta1: float, beta2: float, eps: float, weight_decay: float, step: int,
                 grad_scale: float) -> None:
    """torch.optim.AdamW's update of one parameter tensor in place (GAN_DANet_train.ipynb:182-183); ``step`` is 1-based,
    ``grad_scale`` folds the 1/world of data parallelism."""
    if not (p.is_contiguous() and g.is_contiguous() and m.is_contiguous() and v.is_contiguous()):
        raise L.GdnError("fused_adamw_: contiguous tensors required")
    L.check(E._lib(p).gdn_adamw(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr, beta1, beta2, eps, weight_decay, step, grad_scale,
                                E._stream()), "gdn_adamw")


# ------------------------------------------------------------------------------------------------ BatchNorm statistics
@torch.library.custom_op("gandanet::bn_stats_finalize", mutates_args=("running_mean", "running_var"), device_types="cuda")
def bn_stats_finalize(x: Tensor, weight: Tensor, bias: Tensor, running_mean: Tensor, running_var: Tensor, eps: float, momentum: float) -> Tensor:
    """Train-mode ``nn.BatchNorm2d`` statistics of an NHWC tensor (generator.py:32,62,149,189,219,223): per-channel batch mean / biased variance
    (fp64 accumulation), the momentum update of the running statistics with the UNBIASED variance (in place), and the folded coefficients.
    Returns coef [4, C] = (mean, invstd, scale = weight*invstd, shift = bias - mean*scale): ``y = x*scale + shift``."""
    x = _nhwc(x, "bn_stats_finalize(x)")
    lib = E._lib(x)
    Cc = x.shape[-1]
    M = x.numel() // Cc
    coef = torch.empty((4, Cc), dtype=torch.float32, device=x.device)
    sums = E.colstats(x)
    L.check(lib.gdn_bn_finalize(sums.data_ptr(), M, Cc, weight.data_ptr(), bias.data_ptr(), float(eps), float(momentum), running_mean.data_ptr(), running_var.data_ptr(),
                                coef[0].data_ptr(), coef[1].data_ptr(), coef[2].data_ptr(), coef[3].data_ptr(), E._stream()), "gdn_bn_finalize")
    return coef


@bn_stats_finalize.register_fake
def _(x, weight, bias, running_mean, running_var, eps, momentum):
    return x.new_empty((4, x.shape[-1]))


# ------------------------------------------------------------------------------------------------ bilinear skip fusion
@torch.library.custom_op("gandanet::bilinear_resize_add_fwd", mutates_args=(), device_types="cuda")
def bilinear_resize_add_fwd(s: Tensor, x: Tensor) -> Tensor:
    """``x + F.interpolate(s, size=x.shape[1:3], mode='bilinear', align_corners=False)`` on NHWC tensors: the skip fusion of generator.py:242-246."""
    s, x = _nhwc(s, "bilinear_resize_add_fwd(s)"), _nhwc(x, "bilinear_resize_add_fwd(x)")
    B, Hi, Wi, Cc = s.shape
    _, Ho, Wo, _ = x.shape
    y = x.clone()
    L.check(E._lib(x).gdn_bilinear_fwd(s.data_ptr(), y.data_ptr(), B, Hi, Wi, Ho, Wo, Cc, 1, E._stream()), "gdn_bilinear_fwd")
    return y


@bilinear_resize_add_fwd.register_fake
def _(s, x):
    return torch.empty_like(x)


@torch.library.custom_op("gandanet::bilinear_resize_add_bwd", mutates_args=(), device_types="cuda")
def bilinear_resize_add_bwd(dy: Tensor, hi: int, wi: int) -> Tensor:
    """Gradient of the resized term with respect to ``s`` (the adjoint gather of the bilinear resize); the gradient with respect to ``x`` is dy."""
    dy = _nhwc(dy, "bilinear_resize_add_bwd(dy)")
    B, Ho, Wo, Cc = dy.shape
    ds = torch.empty((B, hi, wi, Cc), dtype=torch.float32, device=dy.device)
    L.check(E._lib(dy).gdn_bilinear_bwd(dy.data_ptr(), ds.data_ptr(), B, hi, wi, Ho, Wo, Cc, 0, E._stream()), "gdn_bilinear_bwd")
    return ds


@bilinear_resize_add_bwd.register_fake
def _(dy, hi, wi):
    return dy.new_empty((dy.shape[0], hi, wi, dy.shape[3]))


def _bil_setup(ctx, inputs, output):
    ctx.hw = (inputs[0].shape[1], inputs[0].shape[2])


def _bil_backward(ctx, dy):
    dy = dy.contiguous()
    return torch.ops.gandanet.bilinear_resize_add_bwd(dy, ctx.hw[0], ctx.hw[1]), dy


bilinear_resize_add_fwd.register_autograd(_bil_backward, setup_context=_bil_setup)


# ------------------------------------------------------------------------------------------------ generator objective
@torch.library.custom_op("gandanet::multi_loss_fwd_bwd", mutates_args=(), device_types="cuda")
def multi_loss_fwd_bwd(hr: Tensor, real: Tensor, fake_logits: Tensor, w_adv: float, tv_weight: float) -> Tuple[Tensor, Tensor, Tensor]:
    """The D-dependent and pixel-space terms of the generator objective in one op (GAN_DANet_train.ipynb:261-267, without the perceptual term):
    ``loss = (1 - w_adv) * MSE(hr, real) + w_adv * BCEWithLogits(fake_logits, 1) + TV(hr; tv_weight)`` (losses.py:76-87 for TV), together with
    its gradients: returns (losses [4] = total, pixel, adv, tv; d loss / d hr [B,1,H,W]; d loss / d fake_logits).  Every term adds value and
    gradient in a single pass over its operand (gdn_mse / gdn_tv accumulate into one gradient field)."""
    if hr.dim() != 4 or hr.shape != real.shape or hr.dtype != torch.float32:
        raise L.GdnError("multi_loss_fwd_bwd: hr and real must be float32 [B, C, H, W] tensors of one shape")
    hr, real, z = hr.contiguous(), real.contiguous(), fake_logits.contiguous()
    lib = E._lib(hr)
    B, Cc, H, W = hr.shape
    dev = hr.device
    part = torch.zeros(3, dtype=torch.float32, device=dev)                     # pixel, adv, tv
    dhr = torch.empty_like(hr)
    dz = torch.empty_like(z)
    ws = E.dot_ws(dev)
    st = E._stream()
    L.check(lib.gdn_mse(hr.data_ptr(), real.data_ptr(), hr.numel(), part[0:1].data_ptr(), dhr.data_ptr(), 1.0 - w_adv, 0, ws.data_ptr(), st), "gdn_mse")
    L.check(lib.gdn_tv(hr.data_ptr(), B * Cc, H, W, float(tv_weight) * Cc, part[2:3].data_ptr(), dhr.data_ptr(), 1.0, 1, ws.data_ptr(), st), "gdn_tv")
    L.check(lib.gdn_bce_logits(z.data_ptr(), z.numel(), None, 1.0, part[1:2].data_ptr(), dz.data_ptr(), float(w_adv), st), "gdn_bce_logits")
    total = ((1.0 - w_adv) * part[0] + w_adv * part[1] + part[2]).reshape(1)
    return torch.cat([total, part]), dhr, dz


@multi_loss_fwd_bwd.register_fake
def _(hr, real, fake_logits, w_adv, tv_weight):
    return hr.new_empty((4,)), torch.empty_like(hr), torch.empty_like(fake_logits)


OPS = ("pam_fwd", "pam_bwd", "cam_fwd", "cam_bwd", "conv2d", "conv2d_bwd", "upsample_bicubic2x", "upsample_bicubic2x_bwd", "fused_adamw_",
       "bn_stats_finalize", "bilinear_resize_add_fwd", "bilinear_resize_add_bwd", "multi_loss_fwd_bwd")
