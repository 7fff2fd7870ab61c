"""Loss functions of GAN-DANet on the B200 kernels (reference ``models/losses.py:13-150``).

``TVLoss``, ``SSIM`` and ``PerceptualLoss`` keep the reference constructor signatures.  ``MSELoss`` and
``BCEWithLogitsLoss`` are kernel-backed stand-ins for the ``torch.nn`` losses the training notebook builds at
``GAN_DANet_train.ipynb:190-191``.  Every loss evaluates value and input-gradient in a single pass; autograd's upstream
scalar is applied with ``gdn_scale_dev`` (device scalar, no host sync).
"""
from __future__ import annotations

import warnings
from typing import Dict, List, Optional, Sequence, Set, Tuple

import itertools
import weakref

import torch
from torch import nn

from .. import _lib as L
from .. import engine as E
from .._lib import ACT_NONE, ACT_RELU


def _scale_by(g: torch.Tensor, scalar: torch.Tensor) -> torch.Tensor:
    out = torch.empty_like(g)
    s = scalar.reshape(1).to(torch.float32).contiguous()
    L.check(E._lib(g).gdn_scale_dev(g.data_ptr(), s.data_ptr(), out.data_ptr(), g.numel(), 0, E._stream()), "gdn_scale_dev")
    return out


class _PairLoss(torch.autograd.Function):
    """mean((a-b)^2) (mode 0) or mean|a-b| (mode 1); gradient w.r.t. both arguments."""

    @staticmethod
    def forward(ctx, a: torch.Tensor, b: torch.Tensor, mode: int):
        a_c, b_c = a.detach().contiguous(), b.detach().contiguous()
        lib = E._lib(a_c)
        loss = torch.empty(1, dtype=torch.float32, device=a.device)
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        grad = torch.empty_like(a_c) if need else None
        ws = E.dot_ws(a.device)
        if mode == 0:
            L.check(lib.gdn_mse(a_c.data_ptr(), b_c.data_ptr(), a_c.numel(), loss.data_ptr(), E._ptr(grad), 1.0, 0, ws.data_ptr(), E._stream()), "gdn_mse")
        else:
            L.check(lib.gdn_l1(a_c.data_ptr(), b_c.data_ptr(), a_c.numel(), loss.data_ptr(), 0, E._ptr(grad), 1.0, 0, ws.data_ptr(), E._stream()), "gdn_l1")
        ctx.grad = grad
        return loss.reshape(())

    @staticmethod
    def backward(ctx, dout):
        ga = gb = None
        if ctx.needs_input_grad[0]:
            ga = _scale_by(ctx.grad, dout)
        if ctx.needs_input_grad[1]:
            gb = _scale_by(ctx.grad, -dout)
        return ga, gb, None


class MSELoss(nn.Module):
    """nn.MSELoss() (mean), GAN_DANet_train.ipynb:191,262."""

    def forward(self, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        return _PairLoss.apply(a, b, 0)


class L1Loss(nn.Module):
    def forward(self, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        return _PairLoss.apply(a, b, 1)


class _BCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z: torch.Tensor, target: torch.Tensor):
        zc = z.detach().contiguous()
        tc = target.detach().to(torch.float32).expand_as(zc).contiguous()
        loss = torch.empty(1, dtype=torch.float32, device=z.device)
        grad = torch.empty_like(zc) if ctx.needs_input_grad[0] else None
        L.check(E._lib(zc).gdn_bce_logits(zc.data_ptr(), zc.numel(), tc.data_ptr(), 0.0, loss.data_ptr(), E._ptr(grad), 1.0, E._stream()), "gdn_bce_logits")
        ctx.grad = grad
        return loss.reshape(())

    @staticmethod
    def backward(ctx, dout):
        return (_scale_by(ctx.grad, dout) if ctx.needs_input_grad[0] else None), None


class BCEWithLogitsLoss(nn.Module):
    """nn.BCEWithLogitsLoss() (mean), GAN_DANet_train.ipynb:190,252-253,261."""

    def forward(self, z: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        return _BCEFn.apply(z, target)


class _TVFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: torch.Tensor, weight: float):
        xc = x.detach().contiguous()
        B, Cc, H, W = xc.shape
        loss = torch.empty(1, dtype=torch.float32, device=x.device)
        grad = torch.empty_like(xc) if ctx.needs_input_grad[0] else None
        # the reference divides by the batch size only (losses.py:87); the kernel treats B*C planes as its batch
        L.check(E._lib(xc).gdn_tv(xc.data_ptr(), B * Cc, H, W, float(weight) * Cc, loss.data_ptr(), E._ptr(grad), 1.0, 0,
                                  E.dot_ws(x.device).data_ptr(), E._stream()), "gdn_tv")
        ctx.grad = grad
        return loss.reshape(())

    @staticmethod
    def backward(ctx, dout):
        return (_scale_by(ctx.grad, dout) if ctx.needs_input_grad[0] else None), None


class TVLoss(nn.Module):
    """Reference losses.py:76-87."""

    def __init__(self, weight: float = 1.0) -> None:
        super().__init__()
        self.weight = weight

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return _TVFn.apply(x, self.weight)


class SSIM(nn.Module):
    """Reference losses.py:90-147 (single-scale, 11x11 Gaussian window, sigma 1.5).  Forward only: the training loop
    evaluates it without adding it to the objective (GAN_DANet_train.ipynb:263 vs :267), so the result is detached."""

    def __init__(self, window_size: int = 11, size_average: bool = True) -> None:
        super().__init__()
        if window_size != 11:
            raise NotImplementedError("the SSIM kernel is specialised for the reference's 11x11 window")
        self.window_size = window_size
        self.size_average = size_average
        self.channel = 1
        self.register_buffer("window", self._create_window(window_size, self.channel))

    def _gaussian(self, window_size: int, sigma: float) -> torch.Tensor:
        coords = torch.arange(window_size, dtype=torch.float32)
        gauss = torch.exp(-((coords - window_size // 2) ** 2) / (2 * sigma ** 2))
        return (gauss / gauss.sum()).unsqueeze(1)

    def _create_window(self, window_size: int, channel: int) -> torch.Tensor:
        _1d = self._gaussian(window_size, 1.5)
        _2d = _1d @ _1d.t()
        window = _2d.float().unsqueeze(0).unsqueeze(0)
        return window.expand(channel, 1, window_size, window_size).contiguous()

    def forward(self, img1: torch.Tensor, img2: torch.Tensor) -> torch.Tensor:
        a, b = img1.detach().contiguous(), img2.detach().contiguous()
        B, Cc, H, W = a.shape
        lib = E._lib(a)
        ws = E.workspace("dot", lib.gdn_ssim_ws_bytes(B * Cc, H, W), a.device)
        if self.size_average:
            out = torch.empty(1, dtype=torch.float32, device=a.device)
            L.check(lib.gdn_ssim(a.data_ptr(), b.data_ptr(), B * Cc, H, W, out.data_ptr(), ws.data_ptr(), E._stream()), "gdn_ssim")
            return out.reshape(())
        out = torch.empty(B, dtype=torch.float32, device=a.device)
        for i in range(B):
            L.check(lib.gdn_ssim(a[i].data_ptr(), b[i].data_ptr(), Cc, H, W, out[i:i + 1].data_ptr(), ws.data_ptr(), E._stream()), "gdn_ssim")
        return out


# torchvision vgg19.features: conv indices and max-pool indices
_VGG_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512, "M"]


def _make_vgg19_features() -> nn.Sequential:
    """Architecture of torchvision.models.vgg19().features (conv3x3+ReLU(inplace) / MaxPool2d(2,2)), default-initialised
    exactly as torchvision does (kaiming_normal_ fan_out, zero bias) when torchvision itself is unavailable."""
    layers: List[nn.Module] = []
    cin = 3
    for v in _VGG_CFG:
        if v == "M":
            layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
        else:
            conv = nn.Conv2d(cin, int(v), kernel_size=3, padding=1)
            nn.init.kaiming_normal_(conv.weight, mode="fan_out", nonlinearity="relu")
            nn.init.constant_(conv.bias, 0)
            layers += [conv, nn.ReLU(inplace=True)]
            cin = int(v)
    return nn.Sequential(*layers)


_UIDS = itertools.count(1)


class PerceptualLoss(nn.Module):
    """VGG-based perceptual loss with optional offline weights (reference losses.py:13-73).

    sum over ``feature_layers`` of mean|phi_i(x3) - phi_i(y3)| with phi = vgg19.features, inputs repeated to three
    channels, no ImageNet normalisation.  VGG is frozen: only the data gradient w.r.t. ``x`` is computed."""

    def __init__(self, feature_layers: Sequence[int] = (1, 6, 11, 20), weights_path: Optional[str] = None, pretrained: bool = True,
                 device: Optional[torch.device] = None, use_gpu: Optional[bool] = None) -> None:
        super().__init__()
        # identity of this instance in the engine's packed-weight cache: a process-wide counter, never id() (CPython reuses addresses of freed
        # objects, and the caching allocator reuses data_ptr()s -- a later member's random VGG could hit a dead member's packed weights)
        self._uid = next(_UIDS)
        weakref.finalize(self, E.drop_frozen_weights, self._uid)
        self.feature_layers: Set[int] = set(feature_layers)
        if not self.feature_layers:
            raise ValueError("feature_layers must contain at least one index")
        max_layer = max(self.feature_layers)
        if device is None:
            # ``use_gpu=`` is what GAN_DANet_train.ipynb:194 passes (the reference signature rejects it)
            want_gpu = torch.cuda.is_available() if use_gpu is None else bool(use_gpu)
            device = torch.device("cuda" if want_gpu and torch.cuda.is_available() else "cpu")
        self.device = torch.device(device)

        vgg_features = None
        if weights_path is None and pretrained:
            try:
                from torchvision import models
                vgg_features = models.vgg19(weights=models.VGG19_Weights.DEFAULT).features
            except Exception:
                warnings.warn("Falling back to randomly initialised VGG19 features. "
                              "Pass pretrained=False or provide weights_path to silence this warning.", RuntimeWarning)
        if vgg_features is None:
            try:
                from torchvision import models
                vgg_features = models.vgg19(weights=None).features
            except Exception:
                vgg_features = _make_vgg19_features()
        if weights_path is not None:
            state_dict = torch.load(weights_path, map_location=self.device)
            missing, unexpected = vgg_features.load_state_dict(state_dict, strict=False)
            if unexpected:
                warnings.warn(f"Unexpected keys when loading VGG weights: {unexpected}", RuntimeWarning)
            if missing:
                warnings.warn(f"Missing keys when loading VGG weights: {missing}", RuntimeWarning)
        self.vgg = nn.Sequential(*list(vgg_features)[: max_layer + 1]).to(self.device)
        self.vgg.eval()
        for param in self.vgg.parameters():
            param.requires_grad_(False)
        self._wcache: Dict[Tuple, Tuple[torch.Tensor, Tuple]] = {}

    def _weights(self, idx: int, conv: nn.Conv2d, cin: int) -> Tuple[torch.Tensor, Tuple]:
        """(OIHW weight, cache key) of a frozen conv; a 1-channel input into conv1_1 uses the channel-summed weight
        (x.repeat(1,3,1,1) == 1-channel conv with sum_c W[:, c], SURVEY appendix A identity 5).  The engine caches the
        kernel operands (permuted fp32 or packed bf16) under the key."""
        key = (self._uid, idx, cin, conv.weight._version)
        hit = self._wcache.get(key)
        if hit is None:
            w = conv.weight.detach()
            if cin != w.shape[1]:
                w = w.sum(dim=1, keepdim=True)      # frozen weights: done once, cached
            hit = (w.contiguous(), key)
            self._wcache[key] = hit
        return hit

    def _bf16_plan(self):
        """Layer structure check for the bf16 feature-map path: every conv is followed by a ReLU and every tap is such a ReLU.
        Returns (conv indices whose output must also exist in fp32, index of the last conv) or None."""
        layers = list(self.vgg)
        convs = [i for i, l in enumerate(layers) if isinstance(l, nn.Conv2d)]
        if not convs or any(i + 1 >= len(layers) or not isinstance(layers[i + 1], nn.ReLU) for i in convs):
            return None
        if any(not (t >= 1 and isinstance(layers[t], nn.ReLU) and isinstance(layers[t - 1], nn.Conv2d)) for t in self.feature_layers):
            return None
        if any(not isinstance(l, (nn.Conv2d, nn.ReLU, nn.MaxPool2d)) for l in layers):
            return None
        return {t - 1 for t in self.feature_layers}, convs[-1]

    def _features_bf16(self, tape: E.Tape, x: E.Var, on_feature, plan, tap1: Optional[dict] = None) -> None:
        """Same features with bf16-only storage of the untapped maps (conv precision 'bf16': the next conv would round them to bf16
        anyway): every conv epilogue writes the next conv's packed operand, max-pool runs on bf16, only tapped maps exist in fp32.
        ``tap1`` (see _tap1_recompute_ok / _tap1_pair): relu1_1 is tapped but never stored -- ``tap1['y16']`` is this branch's conv1_1 + ReLU output as
        the bf16 operand of conv1_2, ``tap1['mask']`` (generated branch) the 2-bit gate / L1-sign codes its backward pass needs."""
        fp32_convs, last_conv = plan
        layers = list(self.vgg)
        cur, cur16 = x, None
        for idx, layer in enumerate(layers):
            if idx == 0 and tap1 is not None:
                cur, cur16 = _frozen_conv1_tap(tape, x, self._weights(0, layer, 1)[0], tap1["y16"], tap1.get("mask"))
                continue
            if idx == 1 and tap1 is not None:
                continue            # the tap's L1 term was accumulated by the pair kernel (PerceptualLoss._tap1_pair)
            if isinstance(layer, nn.Conv2d):
                # a recorded (gradient-carrying) max-pool routes the gradient to the arg-max of the fp32 map: bf16 rounding creates ties
                # that "first maximum in scan order" would break differently (measured: 4 % change of dL/dx), so the maps feeding a
                # pool stay fp32 on the generated branch; the forward-only target branch needs only max VALUES (max and rounding commute)
                feeds_pool = idx + 2 < len(layers) and isinstance(layers[idx + 2], nn.MaxPool2d)
                keep32 = idx in fp32_convs or (tape.record and feeds_pool)
                cur, cur16 = _frozen_conv16(tape, cur, cur16, self._weights(idx, layer, cur.t.shape[-1]), layer.bias.detach(),
                                            want_fp32=keep32, want_16=idx != last_conv and not (tape.record and feeds_pool))
            elif isinstance(layer, nn.MaxPool2d):
                if cur16 is not None:
                    cur, cur16 = E.op_maxpool2_bf16(tape, cur, cur16)
                elif tape.record and idx >= 2 and isinstance(layers[idx - 1], nn.ReLU) and isinstance(layers[idx - 2], nn.Conv2d) and (idx - 1) not in self.feature_layers:
                    cur = E.op_maxpool2_relu_packed(tape, cur)      # untapped conv + ReLU output feeding only this pool: fused gradient route
                else:
                    cur = E.op_maxpool2(tape, cur)
            if idx in self.feature_layers:
                on_feature(idx, cur)

    def _tap1_recompute_ok(self, x: torch.Tensor) -> bool:
        """relu1_1 tapped, single-channel input, bf16 feature-map path, a second conv behind conv1_1 (its data gradient is what the tap's
        gradient is added to): the tapped map can be recomputed instead of stored (engine flag ``vgg_tap1_recompute``)."""
        if not (E.vgg_tap1_recompute and E.bf16_storage_ok() and E.thin_conv_enabled and 1 in self.feature_layers and x.dim() == 4 and x.shape[1] == 1):
            return False
        plan = self._bf16_plan_for(x.shape[2], x.shape[3], 1)
        layers = list(self.vgg)
        if plan is None or plan[1] == 0 or not isinstance(layers[0], nn.Conv2d) or layers[0].kernel_size != (3, 3) or layers[0].padding != (1, 1) \
                or layers[0].stride != (1, 1) or layers[0].bias is None or not isinstance(layers[2], nn.Conv2d):
            return False
        return bool(L.load().gdn_thin_conv_tap_supported(layers[0].out_channels, x.shape[2], x.shape[3]))

    def _bf16_plan_for(self, H: int, W: int, cin0: int):
        """_bf16_plan() if the bf16 feature-map path applies to an input of this grid and channel count, else None."""
        plan = self._bf16_plan() if E.bf16_storage_ok() else None
        if plan is not None:
            # every conv after the first must take the tensor-core path at its resolution (bf16-only maps have no fp32 fallback)
            cin = cin0
            for layer in self.vgg:
                if isinstance(layer, nn.MaxPool2d):
                    H, W = H // 2, W // 2
                elif isinstance(layer, nn.Conv2d):
                    if cin != 1 and not (E.tc_eligible(cin, layer.out_channels, 3, 3, 1, H, W) and layer.out_channels % 8 == 0 and H % 2 == 0 and W % 2 == 0):
                        return None
                    cin = layer.out_channels
            if cin == cin0:
                return None
        return plan

    def _features(self, tape: E.Tape, x: E.Var, on_feature, tap1: Optional[dict] = None) -> None:
        plan = self._bf16_plan_for(x.t.shape[1], x.t.shape[2], x.t.shape[3])
        if plan is not None:
            return self._features_bf16(tape, x, on_feature, plan, tap1)
        assert tap1 is None, "relu1_1 recomputation needs the bf16 feature-map path"
        cur = x
        for idx, layer in enumerate(self.vgg):
            if isinstance(layer, nn.Conv2d):
                fuse_relu = idx + 1 < len(self.vgg) and isinstance(self.vgg[idx + 1], nn.ReLU) and idx not in self.feature_layers
                cur = _frozen_conv(tape, cur, self._weights(idx, layer, cur.t.shape[-1]), layer.bias.detach(), relu=fuse_relu)
                cur_fused = fuse_relu
            elif isinstance(layer, nn.ReLU):
                if not cur_fused:
                    cur = _relu(tape, cur)
                cur_fused = False
            elif isinstance(layer, nn.MaxPool2d):
                cur = E.op_maxpool2(tape, cur)
            else:
                raise NotImplementedError(type(layer))
            if idx in self.feature_layers:
                on_feature(idx, cur)

    def _tap1_pair(self, xin: torch.Tensor, yin: torch.Tensor, loss: torch.Tensor, need_mask: bool):
        """conv1_1 + ReLU of both branches in ONE pass over the two single-channel images (gdn_thin_conv_tap_pair): the bf16 operands of conv1_2
        (generated, target), the relu1_1 L1 term added to ``loss``, and the 2-bit gate / sign codes for the backward pass."""
        conv = self.vgg[0]
        w = self._weights(0, conv, 1)[0]
        B, H, W, _ = xin.shape
        O = w.shape[0]
        dev = xin.device
        lib = E._lib(xin)
        y16a = torch.empty((B * H * W, O), dtype=torch.bfloat16, device=dev)
        y16b = torch.empty((B * H * W, O), dtype=torch.bfloat16, device=dev)
        mask = torch.empty(lib.gdn_thin_conv_tap_mask_bytes(B, H, W, O), dtype=torch.uint8, device=dev) if need_mask else None
        ws = E.dot_ws(dev)
        L.check(lib.gdn_thin_conv_tap_pair(xin.data_ptr(), yin.data_ptr(), w.contiguous().data_ptr(), conv.bias.detach().data_ptr(), B, H, W, O, loss.data_ptr(), 1, 1.0,
                                           y16a.data_ptr(), y16b.data_ptr(), E._ptr(mask), ws.data_ptr(), ws.numel(), E._stream()), "gdn_thin_conv_tap_pair")
        return y16a, y16b, mask

    def _target_features(self, y: torch.Tensor, tap1: Optional[dict] = None, yin: Optional[E.Var] = None):
        """Forward-only features of the target branch: (NHWC input Var, {tap index: tensor})."""
        ytape = E.Tape(record=False)
        yfeat: Dict[int, torch.Tensor] = {}
        if yin is None:
            yin = E.op_from_nchw(ytape, y.detach(), False)
        self._features(ytape, yin, lambda i, f: yfeat.__setitem__(i, f.t), tap1=tap1)
        return yin, yfeat

    def prefetch_target(self, y: torch.Tensor, stream: "torch.cuda.Stream") -> None:
        """Evaluates the target branch (it depends on ``y`` alone: the real field of GAN_DANet_train.ipynb:265) on ``stream``, forked from the
        current stream, so that it runs beside whatever is enqueued next (the trainer: the generator's forward pass).  The next ``forward(x, y)``
        with this ``y`` joins the stream and uses the result.  Not used when relu1_1 is evaluated by the pair kernel (which needs both images).
        The cached tensors live in ``stream``'s allocator pool: they are released before the next prefetch forks from the consumer stream again,
        so a recycled block is never written while the consumer still reads it."""
        self._prefetched = None
        if self._tap1_recompute_ok(y):
            return
        cur = torch.cuda.current_stream(y.device)
        stream.wait_stream(cur)
        with torch.cuda.stream(stream):
            out = self._target_features(y)
            done = torch.cuda.Event()
            done.record(stream)       # the consumer waits for THIS point only, not for what the caller enqueues on ``stream`` afterwards
        self._prefetched = (y.data_ptr(), y._version, tuple(y.shape), out, done)

    def _take_target(self, y: torch.Tensor):
        pre, self._prefetched = getattr(self, "_prefetched", None), None
        if pre is not None and pre[:3] == (y.data_ptr(), y._version, tuple(y.shape)):
            torch.cuda.current_stream(y.device).wait_event(pre[4])
            return pre[3]
        return self._target_features(y)

    def forward(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        return _PerceptualFn.apply(self, x, y)


def _frozen_conv(tape: E.Tape, x: E.Var, wk: Tuple[torch.Tensor, Tuple], bias: torch.Tensor, relu: bool) -> E.Var:
    w, key = wk
    O = w.shape[0]
    B, H, W, _ = x.t.shape
    y = E.Var(E.new_nhwc(B, H, W, O, x.t))
    act = ACT_RELU if relu else ACT_NONE
    cctx = E.conv_forward(x.t, w, y.t, pad=1, bias=bias, act=act, frozen_key=key, keep=False)

    def bwd():
        if y.g is None or not x.needs_grad:
            return
        tgt, acc = x.grad_target()
        Cin = x.t.shape[-1]
        if relu and E.conv_backward_tc_only(O, Cin, 3, 3, 1, H, W, H, W, False, True):
            # ReLU backward fused into the packing of the data-gradient GEMM's operand: the fp32 dz is never stored
            E.conv_backward(cctx, None, x.t, w, pad=1, gx=tgt, gx_accumulate=acc, frozen_key=key,
                            dz_packed=E.pack_actgrad(y.g, y.t, ACT_RELU, 0.0), out_hw=(H, W))
            return
        dz = y.g
        if relu:
            dz = torch.empty_like(y.g)
            E.act_bwd(y.g, y.t, dz, ACT_RELU, 0.0)
        E.conv_backward(cctx, dz, x.t, w, pad=1, gx=tgt, gx_accumulate=acc, frozen_key=key)

    tape.push(bwd)
    return y


def _frozen_conv16(tape: E.Tape, x: E.Var, x16, wk: Tuple[torch.Tensor, Tuple], bias: torch.Tensor, want_fp32: bool, want_16: bool):
    """Frozen 3x3 conv + ReLU of the bf16 feature-map path.  x16: the input as bf16 [B*H*W, Cin] (None: pack x.t, or the 1-channel
    thin kernel).  Returns (Var, bf16 copy or None); without want_fp32 the Var's tensor is a shape-only placeholder that is never
    written or read (its gradient is a real fp32 tensor)."""
    w, key = wk
    O = w.shape[0]
    B, H, W, Cin = x.t.shape
    y = E.Var(E.new_nhwc(B, H, W, O, x.t))
    y16 = torch.empty((B * H * W, O), dtype=torch.bfloat16, device=x.t.device) if want_16 else None
    thin = Cin == 1
    if thin:                       # conv1_1 on the channel-summed weight: fp32 thin kernel, optionally emitting the bf16 copy as well
        want_fp32 = True
        cctx = E.conv_forward(x.t, w, y.t, pad=1, bias=bias, act=ACT_RELU, frozen_key=key, keep=False, y16=y16)
    else:
        cctx = E.conv_forward(x.t, w, y.t if want_fp32 else None, pad=1, bias=bias, act=ACT_RELU, frozen_key=key, keep=False,
                              x_packed=E.Packed(x16, None, Cin) if x16 is not None else None, y16=y16)

    def bwd():
        if (y.g is None and y.g16 is None) or not x.needs_grad:
            return
        tgt, acc = x.grad_target()
        if y.g16 is not None:          # the pool's backward already produced dz = route(dy) * relu'(y) as the packed operand (op_maxpool2_relu_packed)
            assert y.g is None and not thin
            E.conv_backward(cctx, None, x.t, w, pad=1, gx=tgt, gx_accumulate=acc, frozen_key=key, dz_packed=y.g16, out_hw=(H, W))
            return
        if thin and E.thin_gated_dgrad_ok(x.t, w, y.g, y.t, 1, tgt, False, False):
            E.thin_gated_dgrad(y.g, y.t, x.t, w, tgt, stride=1, pad=1, accumulate=acc, slope=0.0)      # ReLU backward inside the gradient kernel's loads
            return
        if thin or not E.conv_backward_tc_only(O, Cin, 3, 3, 1, H, W, H, W, False, True):
            dz = torch.empty_like(y.g)
            E.act_bwd(y.g, y.t, dz, ACT_RELU, 0.0)
            E.conv_backward(cctx, dz, x.t, w, pad=1, gx=tgt, gx_accumulate=acc, frozen_key=key)
            return
        dzp = E.pack_actgrad(y.g, y.t, ACT_RELU, 0.0) if want_fp32 else E.pack_actgrad16(y.g, y16, ACT_RELU, 0.0)
        E.conv_backward(cctx, None, x.t, w, pad=1, gx=tgt, gx_accumulate=acc, frozen_key=key, dz_packed=dzp, out_hw=(H, W))

    tape.push(bwd)
    return y, y16


def _frozen_conv1_tap(tape: E.Tape, x: E.Var, w: torch.Tensor, y16: torch.Tensor, mask: Optional[torch.Tensor]):
    """conv1_1 (1 -> C, channel-summed weight) + ReLU of a branch whose relu1_1 tap is not stored: ``y16`` (written by the pair kernel) is the bf16
    operand of conv1_2.  Backward (generated branch): dL/dx = conv_T((dy + d L1 / d relu1_1) * relu'), gate and L1 sign from ``mask``."""
    O = w.shape[0]
    B, H, W, _ = x.t.shape
    y = E.Var(E.new_nhwc(B, H, W, O, x.t))          # shape carrier (never written or read); its gradient is a real fp32 tensor
    wc = w.contiguous()

    def bwd():
        if y.g is None or not x.needs_grad:
            return
        assert mask is not None and y.g16 is None
        tgt, acc = x.grad_target()
        assert tgt.is_contiguous()
        L.check(E._lib(x.t).gdn_thin_conv_tap_dgrad(y.g.data_ptr(), E.pitch_of(y.g), mask.data_ptr(), wc.data_ptr(), 1.0 / float(B * H * W * O), tgt.data_ptr(),
                                                    tgt.data_ptr() if acc else None, B, H, W, O, E._stream()), "gdn_thin_conv_tap_dgrad")

    tape.push(bwd)
    return y, y16


def _relu(tape: E.Tape, x: E.Var) -> E.Var:
    """Stand-alone ReLU (only where a feature is tapped between a conv and its ReLU)."""
    Cc = x.t.shape[-1]
    one = torch.empty(Cc, dtype=torch.float32, device=x.t.device)
    zero = torch.empty(Cc, dtype=torch.float32, device=x.t.device)
    E.fill_(one, 1.0)
    E.fill_(zero, 0.0)
    y = E.Var(torch.empty_like(x.t))
    E.affine_act(x.t, y.t, one, zero, ACT_RELU)

    def bwd():
        if y.g is None or not x.needs_grad:
            return
        dz = torch.empty_like(y.g)
        E.act_bwd(y.g, y.t, dz, ACT_RELU, 0.0)
        x.add_grad(dz)

    tape.push(bwd)
    return y


class _PerceptualFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod: PerceptualLoss, x: torch.Tensor, y: torch.Tensor):
        if ctx.needs_input_grad[2]:
            raise L.GdnError("PerceptualLoss: the target `y` requires grad; only the first argument is differentiated here (the training loop "
                             "passes the real field as `y`, GAN_DANet_train.ipynb:265).  Detach `y`, or swap the arguments.")
        need = ctx.needs_input_grad[1]
        dev = x.device
        lib = E._lib(x.detach())
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        E.fill_(loss, 0.0)
        ws = E.dot_ws(dev)
        # target branch: forward only
        rec = mod._tap1_recompute_ok(x) and x.shape == y.shape
        tape = E.Tape(record=need)
        xin = E.op_from_nchw(tape, x.detach(), need)
        tap_g = None
        if rec:
            # relu1_1 of both branches from ONE pass over the two images: conv1_2's operands, the tap's L1 term, and the backward pass's 2-bit codes
            yin = E.op_from_nchw(E.Tape(record=False), y.detach(), False)
            y16a, y16b, mask = mod._tap1_pair(xin.t, yin.t, loss, need)
            yin, yfeat = mod._target_features(y, tap1={"y16": y16b}, yin=yin)
            tap_g = {"y16": y16a, "mask": mask}
        else:
            yin, yfeat = mod._take_target(y)       # target branch: forward only (possibly prefetched on a side stream)

        # generated branch: recorded; L1 terms add value and d/dfeature in one pass
        def on_feature(i: int, f: E.Var) -> None:
            t = yfeat[i]
            assert f.t.is_contiguous() and t.is_contiguous()
            g = torch.empty_like(f.t) if need else None
            L.check(lib.gdn_l1(f.t.data_ptr(), t.data_ptr(), f.t.numel(), loss.data_ptr(), 1, E._ptr(g), 1.0, 0, ws.data_ptr(), E._stream()), "gdn_l1")
            if need:
                if f.g is None and f.parent is None:
                    # the L1 term's gradient IS the feature's gradient buffer from here on: the data gradient of the next convolution accumulates
                    # into it in its epilogue (res = gx) instead of a separate read-modify-write pass over the tapped map (1 GB at relu1_1).
                    # Same two summands, same single fp32 addition => bit-identical to adding g afterwards.
                    f.g = g
                else:
                    def bwd():
                        f.add_grad(g)
                    tape.push(bwd)

        mod._features(tape, xin, on_feature, tap1=tap_g)
        ctx.tape, ctx.xin = tape, xin
        return loss.reshape(())

    @staticmethod
    def backward(ctx, dout):
        if not ctx.needs_input_grad[1]:
            return None, None, None
        ctx.tape.backward()
        g = ctx.xin.g
        gx = E.to_nchw(g)
        ctx.tape = ctx.xin = None
        return None, _scale_by(gx, dout), None


__all__ = ["PerceptualLoss", "TVLoss", "SSIM"]
