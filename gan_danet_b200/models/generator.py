"""Generator of GAN-DANet on the B200 kernels.

Same public surface as the reference ``models/generator.py`` (``/root/reference/models/generator.py:11-255``): class
names, constructor signatures, sub-module layout and therefore ``state_dict`` keys are identical, so checkpoints
interchange in both directions and ``module.apply(weights_init_normal)`` consumes the RNG in the same order.  The
``nn.Conv2d`` / ``nn.BatchNorm2d`` children are *parameter containers only*: ``forward`` never calls them, it records
the module on a tape of CUDA kernels from ``libgandanet_sm100.so`` (see ``engine.py``).

Differences that are deliberate and exact (SURVEY appendix A identities):
  * the three bias-free 1x1 skip projections are applied BEFORE the bilinear resize and summed at low resolution
    (a bias-free 1x1 conv commutes with bilinear interpolation; the resize is linear) -- reference :243-245;
  * ``softmax(rowmax(E) - E)`` of CAM is evaluated as ``softmax(-E)`` -- reference :135-136;
  * ``'senet'`` / ``'cbam'`` alias to DANet with a warning, which is what the reference intends at :166-171 but cannot
    do (``warnings`` is never imported there).
"""
from __future__ import annotations

import os
import warnings
from typing import Dict, List, Optional

import torch
from torch import nn

from .. import engine as E
from .._lib import ACT_NONE, ACT_RELU, PREC_FP16, PREC_FP16X3, PREC_FP32

# --- plain PyTorch modules the reference exports but never runs on the hot path --------------------


class OriginalRelationshipLearner(nn.Module):
    """Reference generator.py:11-26.  Instantiated by the notebook but never called; kept as a plain nn.Module."""

    def __init__(self, input_channels: int) -> None:
        super().__init__()
        channels = [64, 128, 256, 512, 1024]
        layers: List[nn.Module] = []
        in_channels = input_channels
        for out_channels in channels:
            layers.append(nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1))
            layers.append(nn.ReLU(inplace=True))
            in_channels = out_channels
        self.net = nn.Sequential(*layers)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.net(x)


class SqueezeExcitation(nn.Module):
    """Reference generator.py:70-85 (unused by the DANet generator)."""

    def __init__(self, channels: int, reduction_ratio: int = 16) -> None:
        super().__init__()
        reduced_channels = max(1, channels // reduction_ratio)
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc1 = nn.Conv2d(channels, reduced_channels, kernel_size=1)
        self.relu = nn.ReLU(inplace=True)
        self.fc2 = nn.Conv2d(reduced_channels, channels, kernel_size=1)
        self.sigmoid = nn.Sigmoid()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        attention = self.avg_pool(x)
        attention = self.relu(self.fc1(attention))
        attention = self.sigmoid(self.fc2(attention))
        return x * attention


class CBAMBlock(nn.Module):
    """Reference generator.py:88-101 (unused by the DANet generator)."""

    def __init__(self, channels: int, reduction_ratio: int = 16) -> None:
        super().__init__()
        self.channel_attention = SqueezeExcitation(channels, reduction_ratio)
        self.spatial_attention = nn.Sequential(nn.Conv2d(2, 1, kernel_size=7, padding=3, bias=False), nn.Sigmoid())

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self.channel_attention(x)
        max_out, _ = torch.max(x, dim=1, keepdim=True)
        avg_out = torch.mean(x, dim=1, keepdim=True)
        attention = self.spatial_attention(torch.cat([max_out, avg_out], dim=1))
        return x * attention


# --- tape plumbing ---------------------------------------------------------------------------------------------------


def default_pam_precision() -> str:
    """'fp16x3' = fused tcgen05 flash kernels with fp16 hi+lo split logit operands (default: the mode that meets the 1e-3 parity bar at the
    reference's unscaled logits), 'fp16' = the same kernels with single fp16 logit operands (about 5 % faster, softmax weights off by up to 5 % at
    |logit| ~ 100), 'fp32' = CUDA-core parity engine.  Env: GDN_PAM_PRECISION."""
    p = os.environ.get("GDN_PAM_PRECISION", "fp16x3").lower()
    if p not in E.PAM_PRECISION_NAMES:
        raise ValueError(f"GDN_PAM_PRECISION={p!r}: expected one of {sorted(E.PAM_PRECISION_NAMES)}")
    return p


class BuildCtx:
    """What a module needs while recording itself: the tape and the Parameter -> Var map of this autograd node."""

    def __init__(self, tape: E.Tape, pmap: Dict[int, E.Var]):
        self.tape, self.pmap = tape, pmap

    def v(self, p: Optional[torch.Tensor]) -> Optional[E.Var]:
        return None if p is None else self.pmap[id(p)]

    def bn(self, m: nn.BatchNorm2d) -> E.BNState:
        return E.BNState(self.v(m.weight), self.v(m.bias), m.running_mean, m.running_var, m.num_batches_tracked, m.eps,
                         m.momentum if m.momentum is not None else 0.1)


class TapeModule(nn.Module):
    """nn.Module whose forward is one autograd node recorded on the kernel tape."""

    def _build(self, ctx: BuildCtx, x: E.Var) -> E.Var:   # pragma: no cover - abstract
        raise NotImplementedError

    def _build_root(self, ctx: BuildCtx, x_nchw: torch.Tensor, x_needs_grad: bool):
        xin = E.op_from_nchw(ctx.tape, x_nchw, x_needs_grad)
        return self._build(ctx, xin), xin

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        params = list(self.parameters())

        def build(tape, xt, x_needs_grad, pvars):
            ctx = BuildCtx(tape, {id(p): v for p, v in zip(params, pvars)})
            return self._build_root(ctx, xt, x_needs_grad)

        return E.TapeFunction.apply(build, x, *params)


def _conv(ctx: BuildCtx, x: E.Var, m: nn.Conv2d, act: int = ACT_NONE, slope: float = 0.0, out: Optional[E.Var] = None) -> E.Var:
    return E.op_conv(ctx.tape, x, ctx.v(m.weight), ctx.v(m.bias), stride=m.stride[0], pad=m.padding[0], act=act, slope=slope, out=out)


# --- hot-path modules ------------------------------------------------------------------------------------------------


class DenseLayer(TapeModule):
    """Reference generator.py:29-38: cat([x, conv3x3(relu(bn(x)))])."""

    def __init__(self, in_channels: int, growth_rate: int) -> None:
        super().__init__()
        self.bn = nn.BatchNorm2d(in_channels)
        self.relu = nn.ReLU(inplace=True)
        self.conv = nn.Conv2d(in_channels, growth_rate, kernel_size=3, padding=1)

    def _build_into(self, ctx: BuildCtx, buf: E.Var, cin: int) -> None:
        """buf[..., :cin] holds x; writes the new features into buf[..., cin:cin+growth]."""
        x = buf.slice(0, cin)
        E.op_bn_act_conv(ctx.tape, x, ctx.bn(self.bn), ctx.v(self.conv.weight), ctx.v(self.conv.bias), training=self.bn.training, act=ACT_RELU,
                         stride=self.conv.stride[0], pad=self.conv.padding[0], out=buf.slice(cin, cin + self.conv.out_channels))

    def _build(self, ctx: BuildCtx, x: E.Var) -> E.Var:
        B, H, W, Cc = x.t.shape
        buf = E.Var(E.new_nhwc(B, H, W, Cc + self.conv.out_channels, x.t))
        E.op_copy(ctx.tape, x, buf.slice(0, Cc))
        self._build_into(ctx, buf, Cc)
        return buf


class DenseBlock(TapeModule):
    """Reference generator.py:41-54.  The concat buffer is allocated once; every layer writes its channel slice."""

    def __init__(self, num_layers: int, in_channels: int, growth_rate: int) -> None:
        super().__init__()
        layers: List[nn.Module] = []
        current_channels = in_channels
        for _ in range(num_layers):
            layers.append(DenseLayer(current_channels, growth_rate))
            current_channels += growth_rate
        self.layers = nn.ModuleList(layers)
        self.in_channels, self.out_channels = in_channels, current_channels

    def alloc(self, B: int, H: int, W: int, like: torch.Tensor) -> E.Var:
        return E.Var(E.new_nhwc(B, H, W, self.out_channels, like))

    def _build_into(self, ctx: BuildCtx, buf: E.Var) -> E.Var:
        c = self.in_channels
        for layer in self.layers:
            layer._build_into(ctx, buf, c)
            c += layer.conv.out_channels
        return buf

    def _build(self, ctx: BuildCtx, x: E.Var) -> E.Var:
        B, H, W, Cc = x.t.shape
        buf = self.alloc(B, H, W, x.t)
        E.op_copy(ctx.tape, x, buf.slice(0, Cc))
        return self._build_into(ctx, buf)


class TransitionLayer(TapeModule):
    """Reference generator.py:57-67: BN -> ReLU -> conv1x1 (no pooling)."""

    def __init__(self, in_channels: int, out_channels: int) -> None:
        super().__init__()
        self.layer = nn.Sequential(nn.BatchNorm2d(in_channels), nn.ReLU(inplace=True), nn.Conv2d(in_channels, out_channels, kernel_size=1))

    def _build(self, ctx: BuildCtx, x: E.Var, out: Optional[E.Var] = None) -> E.Var:
        conv = self.layer[2]
        return E.op_bn_act_conv(ctx.tape, x, ctx.bn(self.layer[0]), ctx.v(conv.weight), ctx.v(conv.bias), training=self.layer[0].training, act=ACT_RELU,
                                stride=conv.stride[0], pad=conv.padding[0], out=out)


class PAMModule(TapeModule):
    """Position attention, reference generator.py:104-122.  ``precision``: 'fp16x3' / 'fp16' run the fused tcgen05/TMEM flash
    kernels (fp16 hi+lo split / single fp16 logit operands, fp32 accumulate), 'fp32' the CUDA-core parity engine."""

    def __init__(self, channels: int) -> None:
        super().__init__()
        reduced_channels = max(1, channels // 8)
        self.query = nn.Conv2d(channels, reduced_channels, kernel_size=1)
        self.key = nn.Conv2d(channels, reduced_channels, kernel_size=1)
        self.value = nn.Conv2d(channels, channels, kernel_size=1)
        self.gamma = nn.Parameter(torch.zeros(1))
        self.precision: Optional[str] = None

    def _build(self, ctx: BuildCtx, x: E.Var, out: Optional[E.Var] = None, y16: Optional[torch.Tensor] = None) -> E.Var:
        prec = (self.precision or default_pam_precision())
        B, H, W, Cc = x.t.shape
        if x.packed is None and E.tc_eligible(Cc, self.query.out_channels, 1, 1, 1, H, W):
            x.packed = E.pack_act(x.t)          # one bf16 operand of x for the three projections (and their three weight gradients)
        if E.pam_merge_qk and x.packed is not None:
            # query and key projections as ONE tensor-core convolution C -> 2d (concatenated weights): one forward, one weight gradient and
            # one data gradient (one read-modify-write pass over dL/dx instead of two); q and k are channel slices of its output
            d = self.query.out_channels
            qk = E.op_conv(ctx.tape, x, E.op_cat_rows(ctx.tape, [ctx.v(self.query.weight), ctx.v(self.key.weight)]),
                           E.op_cat_rows(ctx.tape, [ctx.v(self.query.bias), ctx.v(self.key.bias)]))
            q, k = qk.slice(0, d), qk.slice(d, 2 * d)
        else:
            q = _conv(ctx, x, self.query)
            k = _conv(ctx, x, self.key)
        # the value projection's epilogue also emits the bf16 operand of the fused kernel (no separate packing pass over V)
        if prec not in E.PAM_PRECISION_NAMES:
            raise ValueError(f"PAM precision {prec!r}: expected one of {sorted(E.PAM_PRECISION_NAMES)}")
        v16 = E.pam_v16_buffer(x.t) if prec != "fp32" else None
        v = E.op_conv(ctx.tape, x, ctx.v(self.value.weight), ctx.v(self.value.bias), y16=v16)
        return E.op_pam_core(ctx.tape, x, q, k, v, ctx.v(self.gamma), precision=E.PAM_PRECISION_NAMES[prec], out=out, v16=v16, y16=y16)


class CAMModule(TapeModule):
    """Channel attention, reference generator.py:125-139."""

    def __init__(self, channels: int) -> None:
        super().__init__()
        self.gamma = nn.Parameter(torch.zeros(1))

    def _build(self, ctx: BuildCtx, x: E.Var, out: Optional[E.Var] = None, y16: Optional[torch.Tensor] = None) -> E.Var:
        return E.op_cam(ctx.tape, x, ctx.v(self.gamma), out=out, y16=y16)


class DANetAttention(TapeModule):
    """Reference generator.py:142-157: fuse(cat([PAM(x), CAM(x)])) with conv3x3 + BN + ReLU."""

    def __init__(self, channels: int) -> None:
        super().__init__()
        self.position_attention = PAMModule(channels)
        self.channel_attention = CAMModule(channels)
        self.fuse = nn.Sequential(
            nn.Conv2d(channels * 2, channels, kernel_size=3, padding=1, bias=False),
            nn.BatchNorm2d(channels),
            nn.ReLU(inplace=True),
        )

    def _build(self, ctx: BuildCtx, x: E.Var) -> E.Var:
        B, H, W, Cc = x.t.shape
        cat = E.Var(E.new_nhwc(B, H, W, 2 * Cc, x.t))
        conv = self.fuse[0]
        pam_prec = self.position_attention.precision or default_pam_precision()
        if E.danet_cat16_ok(x.t, pam_prec != "fp32", conv.out_channels):
            # cat[PAM, CAM] is read by the fuse convolution alone: both attention kernels write their result straight into its bf16 operand
            # (column blocks [0, C) and [C, 2C)); the fp32 cat tensor is never written (it only carries the shape and, in backward, the gradient)
            cat16 = torch.empty((B * H * W, 2 * Cc), dtype=torch.bfloat16, device=x.t.device)
            cat.packed = E.Packed(cat16, None, 2 * Cc)
            self.position_attention._build(ctx, x, out=cat.slice(0, Cc), y16=cat16[:, :Cc])
            self.channel_attention._build(ctx, x, out=cat.slice(Cc, 2 * Cc), y16=cat16[:, Cc:])
        else:
            self.position_attention._build(ctx, x, out=cat.slice(0, Cc))
            self.channel_attention._build(ctx, x, out=cat.slice(Cc, 2 * Cc))
        assert conv.bias is None
        return E.op_conv_bn_act(ctx.tape, cat, ctx.v(conv.weight), ctx.bn(self.fuse[1]), training=self.fuse[1].training, act=ACT_RELU,
                                stride=conv.stride[0], pad=conv.padding[0])


def _build_attention(attention_type: Optional[str], channels: int) -> Optional[nn.Module]:
    if attention_type is None or attention_type.lower() == "none":
        return None
    attention = attention_type.lower()
    if attention == "danet":
        return DANetAttention(channels)
    if attention in {"senet", "cbam"}:
        warnings.warn(f"Attention type '{attention_type}' currently aliases to 'danet'.", RuntimeWarning)
        return DANetAttention(channels)
    raise ValueError(f"Unsupported attention type: {attention_type}")


class FlexibleUpsamplingModule(TapeModule):
    """Generator used for super-resolution in GAN-DANet (reference generator.py:175-247)."""

    def __init__(self, input_channels: int = 40, growth_rate: int = 24, num_blocks: int = 3, num_layers_per_block: int = 4,
                 attention_type: Optional[str] = "danet") -> None:
        super().__init__()
        self.initial = nn.Sequential(
            nn.Conv2d(input_channels, 64, kernel_size=3, padding=1, bias=False),
            nn.BatchNorm2d(64),
            nn.ReLU(inplace=True),
        )
        self.dense_blocks = nn.ModuleList()
        self.transition_layers = nn.ModuleList()
        self.attention_modules = nn.ModuleList()
        self.feature_channels: List[int] = []

        num_features = 64
        for block_idx in range(num_blocks):
            dense_block = DenseBlock(num_layers_per_block, num_features, growth_rate)
            self.dense_blocks.append(dense_block)
            num_features += num_layers_per_block * growth_rate
            attention = _build_attention(attention_type, num_features)
            self.attention_modules.append(attention)
            self.feature_channels.append(num_features)
            if block_idx != num_blocks - 1:
                transition = TransitionLayer(num_features, num_features // 2)
                self.transition_layers.append(transition)
                num_features //= 2

        self.channel_adjust = nn.ModuleList([nn.Conv2d(ch, 64, kernel_size=1, bias=False) for ch in reversed(self.feature_channels)])
        self.upsample = nn.Sequential(
            nn.Conv2d(num_features, 64, kernel_size=3, padding=1, bias=False),
            nn.BatchNorm2d(64),
            nn.ReLU(inplace=True),
            nn.Upsample(scale_factor=2, mode="bicubic", align_corners=False),
            nn.Conv2d(64, 64, kernel_size=3, padding=1, bias=False),
            nn.BatchNorm2d(64),
            nn.ReLU(inplace=True),
            nn.Upsample(scale_factor=2, mode="bicubic", align_corners=False),
        )
        self.final = nn.Conv2d(64, 1, kernel_size=3, padding=1)

    @property
    def device(self) -> torch.device:
        """deep_ensemble.ipynb:400-402 reads ``model.device`` (which nn.Module lacks)."""
        return next(self.parameters()).device

    def set_pam_precision(self, precision: Optional[str]) -> None:
        for m in self.modules():
            if isinstance(m, PAMModule):
                m.precision = precision

    def _build_nhwc(self, ctx: BuildCtx, xin: E.Var) -> E.Var:
        """Forward on an NHWC input Var (the fused input-prep path feeds this directly).  With engine.generator_forward_x3 the product mode records
        the forward convolutions with hi+lo split operands (their backward closures run later, on the hi parts, like Discriminator1's)."""
        with E.conv_precision_scope(E.generator_forward_precision()):
            return self._build_nhwc_scoped(ctx, xin)

    def _build_nhwc_scoped(self, ctx: BuildCtx, xin: E.Var) -> E.Var:
        tape = ctx.tape
        B, H, W, _ = xin.t.shape
        nblocks = len(self.dense_blocks)
        buf = self.dense_blocks[0].alloc(B, H, W, xin.t)
        conv0, bn0, dst = self.initial[0], self.initial[1], buf.slice(0, self.dense_blocks[0].in_channels)
        if conv0.bias is None:
            E.op_conv_bn_act(tape, xin, ctx.v(conv0.weight), ctx.bn(bn0), training=bn0.training, act=ACT_RELU, stride=conv0.stride[0], pad=conv0.padding[0], out=dst)
        else:
            E.op_bn_act(tape, _conv(ctx, xin, conv0), ctx.bn(bn0), training=bn0.training, act=ACT_RELU, out=dst)
        skips: List[E.Var] = []
        x = buf
        for i in range(nblocks):
            x = self.dense_blocks[i]._build_into(ctx, buf)
            att = self.attention_modules[i]
            if att is not None:
                x = att._build(ctx, x)
            skips.append(x)
            if i < len(self.transition_layers):
                nxt = self.dense_blocks[i + 1]
                buf = nxt.alloc(B, H, W, xin.t)
                self.transition_layers[i]._build(ctx, x, out=buf.slice(0, nxt.in_channels))
                x = buf
        # skip projections (generator.py:242-246) at the low resolution: the 1x1 convolutions commute with the (linear) bilinear resize
        s: Optional[E.Var] = None
        for adjust, feat in zip(self.channel_adjust, reversed(skips)):
            s = E.op_conv_accumulate(tape, feat, ctx.v(adjust.weight), s)
        up = self.upsample
        stages = ((up[0], up[1]), (up[4], up[5]))
        fused_skip = False
        for k, (conv, bn) in enumerate(stages):
            if conv.bias is None:
                x = E.op_conv_bn_act(tape, x, ctx.v(conv.weight), ctx.bn(bn), training=bn.training, act=ACT_RELU, stride=conv.stride[0], pad=conv.padding[0])
            else:
                x = E.op_bn_act(tape, _conv(ctx, x, conv), ctx.bn(bn), training=bn.training, act=ACT_RELU)
            last = k == len(stages) - 1
            if last and self.final.stride[0] == 1 and self.final.padding[0] == 1 and E.head_tap_planes_ok(x, s, ctx.v(self.final.weight)):
                # up2 -> (+ skips) -> final 3x3 conv (64 -> 1) with the channel reduction hoisted in front of the resampling: no 64-channel tensor at 4h x 4w
                return E.op_upsample_skip_final(tape, x, s, ctx.v(self.final.weight), ctx.v(self.final.bias))
            # otherwise the last up-sampling adds the resized skip sum while it writes its output (one pass over the full-resolution tensor)
            if last and s is not None and E.upsample_skip_fusion and x.t.shape[-1] == s.t.shape[-1] and x.t.shape[-1] % 4 == 0:
                x = E.op_bicubic_up2(tape, x, skip=s)
                fused_skip = True
            else:
                x = E.op_bicubic_up2(tape, x)
        if s is not None and not fused_skip:
            x = E.op_bilinear_add_(tape, s, x)
        return _conv(ctx, x, self.final)

    def _build(self, ctx: BuildCtx, x: E.Var) -> E.Var:
        return self._build_nhwc(ctx, x)


__all__ = ["OriginalRelationshipLearner", "FlexibleUpsamplingModule", "SqueezeExcitation", "CBAMBlock"]
