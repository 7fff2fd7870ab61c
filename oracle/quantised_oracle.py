"""Quantisation-aware CPU oracle of the generator for the tensor-core PRODUCT modes: conv precision 'bf16' with PAM 'fp16x3' (``Formats()``: single
bf16 operands everywhere) and the benchmarked variant of it whose FORWARD convolutions run on hi+lo split operands (``Formats.forward_x3()``:
engine.generator_forward_x3; the gradient GEMMs read the bf16 hi parts of the saved operands).

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as gan_danet_oracle.py: only tests/, smoke() and bench.py's CPU
legs may import it).

It is the reference algorithm (``gan_danet_oracle.generator_forward``, i.e. /root/reference/models/generator.py:230-247) with
rounding hooks at EXACTLY the points where the product-mode kernels round, everything else accumulated in float64
(SURVEY 7.4-1).  The CUDA path is asserted against this oracle (tests/test_gpu_quantised.py: as close to it as the oracle is to its own float32
evaluation for the all-bf16 mode; 1e-3 on the output for the split forward); the oracle's own distance to the plain float64 run of the reference is
the *reported* cost of the operand formats (profiles/r02_precision_product_mode_*.json, r02_precision_benchmarked_mode_gx3.json).

Rounding points of the product mode (gan_danet_b200/engine.py, csrc/conv_tc.cu, csrc/pam_tc.cu):

* every tensor-core convolution (all of the generator's except the 64 -> 1 ``final`` conv): input activation and weight rounded
  to bf16, fp32 accumulate; in the backward the incoming gradient dz is rounded to bf16 for both gradient GEMMs, which use the
  same rounded weight (data gradient) and the same rounded input (weight gradient); the bias gradient sums the unrounded dz;
* BatchNorm + ReLU in front of a dense-layer / transition convolution exist only as that bf16 operand (statistics and the
  backward use the fp32 input) -- the same rounding point as above;
* position attention: logits from fp16 hi+lo split operands (22 mantissa bits: treated as exact); softmax weights P rounded
  to bf16, V rounded to bf16, numerator and denominator both from the rounded P; backward: P = exp(S - lse) with the forward's
  lse (= log of the sum of ROUNDED weights), dy scaled by a power of two and rounded to fp16, dP from the bf16-rounded V,
  dS = P*(dP - rowdot) rounded to fp16, q / k / P rounded to fp16 in the three accumulating products; rowdot = sum dy*o in fp32;
* cat[PAM, CAM] exists only as the bf16 operand of the fuse convolution (again the conv-input rounding point);
* the three skip projections (generator.py:242-246) run BEFORE the bilinear resize (they commute in exact arithmetic), so the
  rounding point of their input is the low-resolution skip tensor;
* CAM (bf16 hi+lo split Gram / re-projection, ~fp32), bicubic / bilinear resampling, BatchNorm, the final convolution: fp32 on
  the device, float64 here.

``fmt_*=None`` switches a rounding point off, which is how the contribution of each one is bisected (tools/precision_bisect.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

import gan_danet_oracle as O

Tensor = torch.Tensor


def rnd(t: Tensor, fmt: Optional[str]) -> Tensor:
    """Round a float64 tensor the way the device rounds its fp32 value to ``fmt`` (round to nearest even)."""
    if fmt is None:
        return t
    if fmt == "bf16":
        return t.float().bfloat16().to(t.dtype)
    if fmt == "fp16":
        return t.float().clamp(-65504.0, 65504.0).half().to(t.dtype)
    if fmt == "fp32":
        return t.float().to(t.dtype)
    if fmt == "bf16x3":     # hi + lo split operand (conv precision 'bf16x3'): hi = bf16(v), lo = bf16(v - hi); the lo*lo product is dropped on the device (~2^-17 relative)
        v = t.float()
        hi = v.bfloat16().float()
        return (hi + (v - hi).bfloat16().float()).to(t.dtype)
    raise ValueError(fmt)


@dataclass
class Formats:
    """Which format each rounding point uses (None = exact)."""
    conv_x: Optional[str] = "bf16"       # convolution input operand
    conv_w: Optional[str] = "bf16"       # convolution weight operand
    conv_g: Optional[str] = "bf16"       # gradient operand of the convolution's two gradient GEMMs
    pam_p: Optional[str] = "bf16"        # forward softmax weights
    pam_v: Optional[str] = "bf16"        # value operand (forward, and dP = dy V^T in the backward)
    pam_g: Optional[str] = "fp16"        # backward: scaled dy, dS, q, k, P
    pam_logits: Optional[str] = None     # 'fp16' emulates the single-fp16 logit operands of PAM mode 'fp16'; None = split (exact)
    # formats of the SAVED input / weight operands that the two gradient GEMMs read; "same" = the forward's.  The mixed mode (forward with hi+lo split
    # operands, backward on the hi parts alone: engine.generator_forward_x3, Discriminator1's forward) is conv_x = conv_w = 'bf16x3', *_bwd = 'bf16'
    conv_x_bwd: Optional[str] = "same"
    conv_w_bwd: Optional[str] = "same"

    @staticmethod
    def exact() -> "Formats":
        return Formats(None, None, None, None, None, None, None)

    @staticmethod
    def forward_x3() -> "Formats":
        """Product mode with the generator's forward convolutions on hi+lo split operands (their gradient GEMMs read the hi parts: bf16)."""
        return Formats(conv_x="bf16x3", conv_w="bf16x3", conv_x_bwd="bf16", conv_w_bwd="bf16")


class QConv(torch.autograd.Function):
    """conv2d with rounded operands (forward) and a rounded gradient operand (backward)."""

    @staticmethod
    def forward(ctx, x, w, b, stride, padding, f: Formats):
        xq, wq = rnd(x, f.conv_x), rnd(w, f.conv_w)
        ctx.save_for_backward(xq if f.conv_x_bwd == "same" else rnd(x, f.conv_x_bwd), wq if f.conv_w_bwd == "same" else rnd(w, f.conv_w_bwd))
        ctx.cfg = (stride, padding, f, b is not None)
        return F.conv2d(xq, wq, b, stride=stride, padding=padding)

    @staticmethod
    def backward(ctx, g):
        xq, wq = ctx.saved_tensors
        stride, padding, f, has_b = ctx.cfg
        gq = rnd(g, f.conv_g)
        dx = torch.nn.grad.conv2d_input(xq.shape, wq, gq, stride=stride, padding=padding) if ctx.needs_input_grad[0] else None
        dw = torch.nn.grad.conv2d_weight(xq, wq.shape, gq, stride=stride, padding=padding) if ctx.needs_input_grad[1] else None
        db = g.sum(dim=(0, 2, 3)) if (has_b and ctx.needs_input_grad[2]) else None
        return dx, dw, db, None, None, None


def qconv(x, w, b, f: Formats, stride: int = 1, padding: int = 0):
    return QConv.apply(x, w, b, stride, padding, f)


class QPamCore(torch.autograd.Function):
    """The fused position-attention kernels' arithmetic: y = gamma * softmax(q k^T) v + x on [B, N, *] operands."""

    @staticmethod
    def forward(ctx, x, q, k, v, gamma, f: Formats):
        ql, kl = rnd(q, f.pam_logits), rnd(k, f.pam_logits)
        S = ql @ kl.transpose(1, 2)
        m = S.max(dim=-1, keepdim=True)[0]
        Pt = rnd(torch.exp(S - m), f.pam_p)
        l = Pt.sum(dim=-1, keepdim=True)
        vq = rnd(v, f.pam_v)
        o = (Pt @ vq) / l
        lse = m + torch.log(l)
        ctx.save_for_backward(q, k, vq, o, lse, gamma, ql, kl)
        ctx.f = f
        return gamma * o + x

    @staticmethod
    def backward(ctx, dy):
        q, k, vq, o, lse, gamma, ql, kl = ctx.saved_tensors
        f = ctx.f
        rowdot = (dy * o).sum(dim=-1, keepdim=True)
        amax = float(dy.abs().max())
        if f.pam_g is not None and amax > 0:
            e = math.floor(math.log2(amax))
            sc = 2.0 ** (-(e + 1))
        else:
            sc = 1.0
        dys = rnd(dy * sc, f.pam_g)
        P = torch.exp(ql @ kl.transpose(1, 2) - lse)
        dP = dys @ vq.transpose(1, 2)
        dS = rnd(P * (dP - rowdot * sc), f.pam_g)
        g = gamma / sc
        dq = g * (dS @ rnd(k, f.pam_g))
        dk = g * (dS.transpose(1, 2) @ rnd(q, f.pam_g))
        dv = g * (rnd(P, f.pam_g).transpose(1, 2) @ dys)
        dgamma = rowdot.sum().reshape(gamma.shape)
        return dy, dq, dk, dv, dgamma, None


def qpam(x: Tensor, sd: Dict[str, Tensor], prefix: str, f: Formats) -> Tensor:
    """PAMModule (generator.py:104-122) in the product mode: the three projections are tensor-core 1x1 convolutions."""
    b, c, h, w = x.shape
    n = h * w
    q = qconv(x, sd[prefix + "query.weight"], sd[prefix + "query.bias"], f).reshape(b, -1, n).transpose(1, 2)
    k = qconv(x, sd[prefix + "key.weight"], sd[prefix + "key.bias"], f).reshape(b, -1, n).transpose(1, 2)
    v = qconv(x, sd[prefix + "value.weight"], sd[prefix + "value.bias"], f).reshape(b, -1, n).transpose(1, 2)
    xr = x.reshape(b, c, n).transpose(1, 2)
    y = QPamCore.apply(xr, q, k, v, sd[prefix + "gamma"], f)
    return y.transpose(1, 2).reshape(b, c, h, w)


def generator_forward(sd: Dict[str, Tensor], x: Tensor, f: Formats, training: bool = True, buffers_out: Optional[Dict[str, Tensor]] = None,
                      taps: Optional[Dict[str, Tensor]] = None, fmt_for=None) -> Tensor:
    """``gan_danet_oracle.generator_forward`` with the product mode's rounding points (see the module docstring).
    ``fmt_for(name) -> Formats`` overrides the formats per convolution (names: initial, dense{b}.{l}, pam{b}, fuse{b}, trans{b}, up0, up1, adjust{i}):
    which layers' operand rounding costs what (tools/precision_bisect.py --layers)."""
    blocks, layers, has_attn = O.generator_structure(sd)
    f0 = f

    def fm(name):
        return f0 if fmt_for is None else fmt_for(name)

    def tap(name, t):
        if taps is not None:
            taps[name] = t
        return t

    x = qconv(x, sd["initial.0.weight"], None, fm("initial"), padding=1)
    x = tap("initial", O.relu(O._bn(sd, "initial.1.", x, training, buffers_out)))
    skips: List[Tensor] = []
    for bi in range(blocks):
        for li in range(layers):
            p = f"dense_blocks.{bi}.layers.{li}."
            y = O.relu(O._bn(sd, p + "bn.", x, training, buffers_out))
            y = qconv(y, sd[p + "conv.weight"], sd[p + "conv.bias"], fm(f"dense{bi}.{li}"), padding=1)
            x = torch.cat([x, y], dim=1)
        tap(f"dense{bi}", x)
        if has_attn:
            a = f"attention_modules.{bi}."
            pos = tap(f"pam{bi}", qpam(x, sd, a + "position_attention.", fm(f"pam{bi}")))
            ch = tap(f"cam{bi}", O.cam(x, sd[a + "channel_attention.gamma"]))
            fz = qconv(torch.cat([pos, ch], dim=1), sd[a + "fuse.0.weight"], None, fm(f"fuse{bi}"), padding=1)
            x = O.relu(O._bn(sd, a + "fuse.1.", fz, training, buffers_out))
        skips.append(tap(f"skip{bi}", x))
        if bi != blocks - 1:
            t = f"transition_layers.{bi}.layer."
            y = O.relu(O._bn(sd, t + "0.", x, training, buffers_out))
            x = tap(f"trans{bi}", qconv(y, sd[t + "2.weight"], sd[t + "2.bias"], fm(f"trans{bi}")))
    x = qconv(x, sd["upsample.0.weight"], None, fm("up0"), padding=1)
    x = O.bicubic_up2(O.relu(O._bn(sd, "upsample.1.", x, training, buffers_out)))
    x = qconv(x, sd["upsample.4.weight"], None, fm("up1"), padding=1)
    x = O.bicubic_up2(O.relu(O._bn(sd, "upsample.5.", x, training, buffers_out)))
    tap("up2", x)
    s = None
    for i, feat in enumerate(reversed(skips)):          # projections hoisted in front of the (linear) resize, as on the device
        pz = qconv(feat, sd[f"channel_adjust.{i}.weight"], None, fm(f"adjust{i}"))
        s = pz if s is None else s + pz
    if s is not None:
        x = x + O.bilinear_to(s, (x.shape[2], x.shape[3]))
    tap("fused", x)
    return O.conv2d(x, sd["final.weight"], sd["final.bias"], padding=1)
