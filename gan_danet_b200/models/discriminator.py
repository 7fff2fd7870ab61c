"""Discriminators of GAN-DANet on the B200 kernels (reference ``models/discriminator.py:8-80``).

``Discriminator1`` keeps the reference layout (``conv1..4``, lazy ``fc1``, ``fc2``, ``activation``) and state_dict keys.
``fc1`` is an ``nn.LazyLinear`` exactly as in the reference; it is materialised on the first forward through
LazyLinear's own ``_infer_parameters`` hook, so initialisation and RNG consumption match.
``SRGAND`` (exported by the reference, never instantiated by its notebooks) runs on the same conv / BatchNorm kernels.
"""
from __future__ import annotations

import torch
from torch import nn
from torch.nn.parameter import UninitializedParameter

from .. import engine as E
from .._lib import ACT_LRELU, ACT_NONE
from .generator import BuildCtx, TapeModule, _conv


def _flatten_nchw(ctx: BuildCtx, x: E.Var) -> E.Var:
    """torch.flatten(x, 1) of the reference's NCHW tensor: [B,H,W,C] NHWC -> [B, C*H*W] in (c, h, w) order."""
    B, H, W, Cc = x.t.shape
    y = E.Var(torch.empty((B, Cc * H * W), dtype=torch.float32, device=x.t.device))
    E.nhwc_to_nchw(x.t, y.t.view(B, Cc, H, W))

    def bwd():
        if y.g is None or not x.needs_grad:
            return
        gx = E.new_nhwc(B, H, W, Cc, x.t)
        E.nchw_to_nhwc(y.g.view(B, Cc, H, W), gx)
        x.add_grad(gx)

    ctx.tape.push(bwd)
    return y


class Discriminator1(TapeModule):
    """Lightweight discriminator with lazy linear projection (reference discriminator.py:57-77)."""

    def __init__(self, input_channels: int = 1) -> None:
        super().__init__()
        self.conv1 = nn.Conv2d(input_channels, 64, kernel_size=3, stride=2, padding=1)
        self.conv2 = nn.Conv2d(64, 128, kernel_size=3, stride=2, padding=1)
        self.conv3 = nn.Conv2d(128, 256, kernel_size=3, stride=2, padding=1)
        self.conv4 = nn.Conv2d(256, 512, kernel_size=3, stride=2, padding=1)
        self.fc1 = nn.LazyLinear(1024)
        self.fc2 = nn.Linear(1024, 1)
        self.activation = nn.LeakyReLU(negative_slope=0.2, inplace=True)

    def _materialise_fc1(self, x: torch.Tensor) -> None:
        if isinstance(self.fc1.weight, UninitializedParameter):
            h, w = x.shape[2], x.shape[3]
            for _ in range(4):
                h, w = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
            dummy = torch.empty((1, 512 * h * w), dtype=x.dtype, device=x.device)
            with torch.no_grad():
                self.fc1._infer_parameters(self.fc1, (dummy,))   # LazyLinear -> Linear, default nn.Linear init

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        self._materialise_fc1(x)
        return super().forward(x)

    def _build(self, ctx: BuildCtx, x: E.Var) -> E.Var:
        slope = self.activation.negative_slope
        with E.conv_precision_scope(E.discriminator_forward_precision()):      # product mode: hi+lo split operands in the forward (no normalisation in D)
            for conv in (self.conv1, self.conv2, self.conv3, self.conv4):
                x = _conv(ctx, x, conv, act=ACT_LRELU, slope=slope)
        f = _flatten_nchw(ctx, x)
        h = E.op_linear(ctx.tape, f, ctx.v(self.fc1.weight), ctx.v(self.fc1.bias), act=ACT_LRELU, slope=slope)
        return E.op_linear(ctx.tape, h, ctx.v(self.fc2.weight), ctx.v(self.fc2.bias))


class SRGAND(TapeModule):
    """Patch-based discriminator inspired by SRGAN (reference discriminator.py:8-54)."""

    def __init__(self, dim: int = 64, in_channels: int = 1) -> None:
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, dim, kernel_size=4, stride=2, padding=1)
        self.conv2 = nn.Conv2d(dim, dim * 2, kernel_size=4, stride=2, padding=1)
        self.bn1 = nn.BatchNorm2d(dim * 2)
        self.conv3 = nn.Conv2d(dim * 2, dim * 4, kernel_size=4, stride=2, padding=1)
        self.bn2 = nn.BatchNorm2d(dim * 4)
        self.conv4 = nn.Conv2d(dim * 4, dim * 8, kernel_size=4, stride=2, padding=1)
        self.bn3 = nn.BatchNorm2d(dim * 8)
        self.conv5 = nn.Conv2d(dim * 8, dim * 16, kernel_size=4, stride=2, padding=1)
        self.bn4 = nn.BatchNorm2d(dim * 16)
        self.conv6 = nn.Conv2d(dim * 16, dim * 32, kernel_size=4, stride=2, padding=1)
        self.bn5 = nn.BatchNorm2d(dim * 32)
        self.conv7 = nn.Conv2d(dim * 32, dim * 16, kernel_size=1)
        self.bn6 = nn.BatchNorm2d(dim * 16)
        self.conv8 = nn.Conv2d(dim * 16, dim * 8, kernel_size=1)
        self.bn7 = nn.BatchNorm2d(dim * 8)
        self.conv9 = nn.Conv2d(dim * 8, dim * 2, kernel_size=1)
        self.bn8 = nn.BatchNorm2d(dim * 2)
        self.conv10 = nn.Conv2d(dim * 2, dim * 2, kernel_size=3, padding=1)
        self.bn9 = nn.BatchNorm2d(dim * 2)
        self.conv11 = nn.Conv2d(dim * 2, dim * 8, kernel_size=3, padding=1)
        self.bn10 = nn.BatchNorm2d(dim * 8)
        self.global_avg_pool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(dim * 8, 1)
        self.activation = nn.LeakyReLU(0.2, inplace=True)

    def _build(self, ctx: BuildCtx, x: E.Var) -> E.Var:
        slope = self.activation.negative_slope
        tape = ctx.tape

        def cba(x, conv, bn):
            return E.op_bn_act(tape, _conv(ctx, x, conv), ctx.bn(bn), training=bn.training, act=ACT_LRELU, slope=slope)

        x = _conv(ctx, x, self.conv1, act=ACT_LRELU, slope=slope)
        x = cba(x, self.conv2, self.bn1)
        x = cba(x, self.conv3, self.bn2)
        x = cba(x, self.conv4, self.bn3)
        x = cba(x, self.conv5, self.bn4)
        x = cba(x, self.conv6, self.bn5)
        x = cba(x, self.conv7, self.bn6)
        x = cba(x, self.conv8, self.bn7)
        residual = x
        y = cba(x, self.conv9, self.bn8)
        y = cba(y, self.conv10, self.bn9)
        y = cba(y, self.conv11, self.bn10)
        # x = x + residual ; global average pool ; fc
        E.op_add_(tape, y, residual)
        pooled = E.op_global_avg_pool(tape, y)
        return E.op_linear(tape, pooled, ctx.v(self.fc.weight), ctx.v(self.fc.bias))


__all__ = ["SRGAND", "Discriminator1"]
