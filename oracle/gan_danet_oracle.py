"""CPU oracle for the GAN-DANet generator/discriminator training step.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product path (``gan_danet_b200``) never
does and fails loudly when the CUDA library is missing.

It restates, as plain functional PyTorch-on-CPU code over a ``state_dict``, the
algorithm of the reference modules (file:line are into ``/root/reference``):

* ``models/generator.py:29-67``   dense layer / dense block / transition
* ``models/generator.py:104-122`` position attention (PAM)
* ``models/generator.py:125-139`` channel attention (CAM)
* ``models/generator.py:142-157`` DANet fuse
* ``models/generator.py:175-247`` FlexibleUpsamplingModule
* ``models/discriminator.py:57-77`` Discriminator1
* ``models/losses.py:13-147``     Perceptual / TV / SSIM
* ``GAN_DANet_train.ipynb:182-194,225-269`` the G+D step (AdamW, BCE, MSE, ...)

The arithmetic of the reference lives in a third-party dependency (PyTorch
ATen, unpinned in ``requirement.yml:14-16``; oracle version = the torch in this
image, 2.11.0 CPU kernels).  Every non-trivial ATen behaviour the path relies
on (train-mode BatchNorm incl. running statistics, bicubic A=-0.75 / bilinear
``align_corners=False`` resampling, max-subtracted softmax, BCE-with-logits,
AdamW) is written out explicitly here instead of being called, so the oracle
states the algorithm rather than delegating it.  Convolutions use
``F.conv2d`` (cross-correlation, zero padding) -- the one primitive kept.

Pinning: the reference holds no tests or golden vectors for this path (SURVEY
§4, §8c).  The oracle is pinned against outputs of the *reference itself*
imported in the build container (``oracle/make_golden.py`` -> fixtures under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks the oracle against them and
re-checks live against the reference when ``/root/reference`` is present).

Everything works in float32 or float64 (pass double tensors).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

BN_EPS = 1e-5
BN_MOMENTUM = 0.1

# ----------------------------------------------------------------------------
# primitives
# ----------------------------------------------------------------------------


def conv2d(x: Tensor, w: Tensor, b: Optional[Tensor] = None, stride: int = 1, padding: int = 0) -> Tensor:
    return F.conv2d(x, w, b, stride=stride, padding=padding)


def batchnorm(x: Tensor, weight: Tensor, bias: Tensor, running_mean: Tensor, running_var: Tensor,
              training: bool, buffers_out: Optional[SD] = None, prefix: str = "") -> Tensor:
    """``nn.BatchNorm2d``.  In training mode uses biased batch
    variance to normalise and writes the momentum-0.1 / unbiased-variance running
    statistics into ``buffers_out`` (so the oracle never mutates its inputs)."""
    if training:
        n = x.shape[0] * x.shape[2] * x.shape[3]
        mean = x.mean(dim=(0, 2, 3))
        var = ((x - mean[None, :, None, None]) ** 2).mean(dim=(0, 2, 3))
        if buffers_out is not None:
            with torch.no_grad():
                buffers_out[prefix + "running_mean"] = (1 - BN_MOMENTUM) * running_mean + BN_MOMENTUM * mean
                buffers_out[prefix + "running_var"] = (1 - BN_MOMENTUM) * running_var + BN_MOMENTUM * var * (n / max(n - 1, 1))
    else:
        mean, var = running_mean, running_var
    inv = torch.rsqrt(var + BN_EPS)
    return (x - mean[None, :, None, None]) * (inv * weight)[None, :, None, None] + bias[None, :, None, None]


def relu(x: Tensor) -> Tensor:
    return torch.clamp_min(x, 0)


def leaky_relu(x: Tensor, slope: float = 0.2) -> Tensor:
    return torch.where(x > 0, x, slope * x)


def softmax_lastdim(x: Tensor) -> Tensor:
    m = x.max(dim=-1, keepdim=True)[0]
    e = torch.exp(x - m)
    return e / e.sum(dim=-1, keepdim=True)


def _cubic_weights(t: float, a: float = -0.75) -> List[float]:
    def near(u):   # |u| <= 1
        return ((a + 2) * u - (a + 3)) * u * u + 1

    def far(u):    # 1 < |u| < 2
        return ((a * u - 5 * a) * u + 8 * a) * u - 4 * a

    return [far(t + 1), near(t), near(1 - t), far(2 - t)]


def _resample_matrix(n_in: int, n_out: int, mode: str, dtype, scale: Optional[float] = None) -> Tensor:
    """Dense [n_out, n_in] interpolation matrix of torch's ``align_corners=False``
    resamplers along one axis (source index (dst+0.5)/scale-0.5,
    border-clamped taps, cubic A=-0.75, bilinear source clamped at 0)."""
    m = torch.zeros(n_out, n_in, dtype=torch.float64)
    ratio = (1.0 / scale) if scale is not None else n_in / n_out
    for o in range(n_out):
        s = (o + 0.5) * ratio - 0.5
        if mode == "bicubic":
            i0 = math.floor(s)
            t = s - i0
            for k, wk in enumerate(_cubic_weights(t)):
                idx = min(max(i0 - 1 + k, 0), n_in - 1)
                m[o, idx] += wk
        elif mode == "bilinear":
            s = max(s, 0.0)
            i0 = min(int(math.floor(s)), n_in - 1)
            i1 = min(i0 + 1, n_in - 1)
            t = s - i0
            m[o, i0] += 1 - t
            m[o, i1] += t
        else:
            raise ValueError(mode)
    return m.to(dtype)


def resample2d(x: Tensor, out_hw: Tuple[int, int], mode: str, scale: Optional[float] = None) -> Tensor:
    """Separable resize of NCHW ``x`` to ``out_hw`` (bicubic or bilinear)."""
    mh = _resample_matrix(x.shape[2], out_hw[0], mode, x.dtype, scale)
    mw = _resample_matrix(x.shape[3], out_hw[1], mode, x.dtype, scale)
    y = torch.einsum("oh,bchw->bcow", mh, x)
    return torch.einsum("pw,bcow->bcop", mw, y)


def bicubic_up2(x: Tensor) -> Tensor:
    """``nn.Upsample(scale_factor=2, mode='bicubic')`` (generator.py:221,225)."""
    return resample2d(x, (2 * x.shape[2], 2 * x.shape[3]), "bicubic", scale=2.0)


def bicubic_down(x: Tensor, scale: float) -> Tensor:
    """``F.interpolate(x, scale_factor=0.5|0.25, mode='bicubic')`` -- the input
    preparation of ``GAN_DANet_train.ipynb:226,231`` (no antialiasing)."""
    oh, ow = int(math.floor(x.shape[2] * scale)), int(math.floor(x.shape[3] * scale))
    return resample2d(x, (oh, ow), "bicubic", scale=scale)


def bilinear_to(x: Tensor, out_hw: Tuple[int, int]) -> Tensor:
    """``F.interpolate(x, size=out_hw, mode='bilinear')`` (generator.py:244)."""
    return resample2d(x, out_hw, "bilinear")


# ----------------------------------------------------------------------------
# attention modules
# ----------------------------------------------------------------------------


def pam(x: Tensor, wq: Tensor, bq: Tensor, wk: Tensor, bk: Tensor, wv: Tensor, bv: Tensor, gamma: Tensor) -> Tensor:
    """Position attention, generator.py:104-122.  No 1/sqrt(d) scaling."""
    b, c, h, w = x.shape
    n = h * w
    q = conv2d(x, wq, bq).reshape(b, -1, n)            # [B, d, N]
    k = conv2d(x, wk, bk).reshape(b, -1, n)            # [B, d, N]
    v = conv2d(x, wv, bv).reshape(b, -1, n)            # [B, C, N]
    energy = torch.einsum("bdi,bdj->bij", q, k)        # [B, N, N]
    attn = softmax_lastdim(energy)
    out = torch.einsum("bcj,bij->bci", v, attn).reshape(b, c, h, w)
    return gamma * out + x


def pam_blocked(x: Tensor, wq, bq, wk, bk, wv, bv, gamma, block: int = 1024) -> Tensor:
    """Same mathematics as :func:`pam`, looping over query-row blocks so the NxN
    map never exists (needed for N >~ 30000; SURVEY §8c caveat 5)."""
    b, c, h, w = x.shape
    n = h * w
    q = conv2d(x, wq, bq).reshape(b, -1, n)
    k = conv2d(x, wk, bk).reshape(b, -1, n)
    v = conv2d(x, wv, bv).reshape(b, -1, n)
    outs = []
    vt = v.transpose(1, 2).contiguous()                              # [B, N, C]
    for i0 in range(0, n, block):
        qb = q[:, :, i0:i0 + block].transpose(1, 2).contiguous()      # [B, blk, d]; plain matmuls: einsum on the strided slice is 10x slower
        a = softmax_lastdim(qb @ k)                                   # energy rows i0.., generator.py:115-116
        outs.append((a @ vt).transpose(1, 2))                         # out[b, c, i] = sum_j v[b, c, j] attn[b, i, j], :120
    out = torch.cat(outs, dim=2).reshape(b, c, h, w)
    return gamma * out + x


def pam_rows(x: Tensor, wq, bq, wk, bk, wv, bv, gamma, rows: Tensor) -> Tensor:
    """Rows ``rows`` (flat positions i = y*W + x) of :func:`pam`'s output: [B, C, len(rows)].  The same mathematics as
    generator.py:113-122 restricted to those query positions (every key/value position still enters their softmax), so
    the oracle of the high-resolution variant (N = 51200, BASELINE configs[3]) costs len(rows)*N instead of N*N.
    Differentiable: a loss that only reads these rows has exactly the gradients of the full module under a cotangent
    that is zero elsewhere."""
    b, c, h, w = x.shape
    n = h * w
    q = conv2d(x, wq, bq).reshape(b, -1, n)[:, :, rows]
    k = conv2d(x, wk, bk).reshape(b, -1, n)
    v = conv2d(x, wv, bv).reshape(b, -1, n)
    attn = softmax_lastdim(torch.einsum("bdi,bdj->bij", q, k))
    out = torch.einsum("bcj,bij->bci", v, attn)
    return gamma * out + x.reshape(b, c, n)[:, :, rows]


def pam_core_flash(q: Tensor, k: Tensor, v: Tensor, block: int = 128) -> Tuple[Tensor, Tensor]:
    """Online-softmax restatement of the PAM core on row-major operands
    (q,k: [B,N,d]; v: [B,N,C]) -> (o [B,N,C], lse [B,N]).  This is the
    formulation the CUDA kernels implement."""
    bsz, n, _ = q.shape
    o = torch.zeros(bsz, n, v.shape[2], dtype=q.dtype)
    m = torch.full((bsz, n), -float("inf"), dtype=q.dtype)
    l = torch.zeros(bsz, n, dtype=q.dtype)
    for j0 in range(0, n, block):
        s = q @ k[:, j0:j0 + block].transpose(1, 2)
        m_new = torch.maximum(m, s.max(dim=-1)[0])
        alpha = torch.exp(m - m_new)
        p = torch.exp(s - m_new[..., None])
        l = l * alpha + p.sum(dim=-1)
        o = o * alpha[..., None] + p @ v[:, j0:j0 + block]
        m = m_new
    return o / l[..., None], m + torch.log(l)


def pam_core_backward(q: Tensor, k: Tensor, v: Tensor, o: Tensor, lse: Tensor, do: Tensor,
                      block: int = 128) -> Tuple[Tensor, Tensor, Tensor]:
    """Flash-style backward of the PAM core:
    P = exp(S - lse), dV = P^T dO, dP = dO V^T, dS = P*(dP - delta), dQ = dS K, dK = dS^T Q."""
    n = q.shape[1]
    delta = (do * o).sum(dim=-1)                                    # [B,N]
    dq = torch.zeros_like(q)
    dk = torch.zeros_like(k)
    dv = torch.zeros_like(v)
    for j0 in range(0, n, block):
        kj, vj = k[:, j0:j0 + block], v[:, j0:j0 + block]
        p = torch.exp(q @ kj.transpose(1, 2) - lse[..., None])     # [B,N,bj]
        dv[:, j0:j0 + block] = p.transpose(1, 2) @ do
        ds = p * (do @ vj.transpose(1, 2) - delta[..., None])
        dq += ds @ kj
        dk[:, j0:j0 + block] = ds.transpose(1, 2) @ q
    return dq, dk, dv


def cam(x: Tensor, gamma: Tensor) -> Tensor:
    """Channel attention, generator.py:125-139 (softmax(rowmax(E)-E))."""
    b, c, h, w = x.shape
    xm = x.reshape(b, c, -1)
    energy = torch.einsum("bin,bjn->bij", xm, xm)
    energy_new = energy.max(dim=-1, keepdim=True)[0] - energy
    attn = softmax_lastdim(energy_new)
    out = torch.einsum("bij,bjn->bin", attn, xm).reshape(b, c, h, w)
    return gamma * out + x


def cam_backward(x: Tensor, gamma: Tensor, dy: Tensor) -> Tuple[Tensor, Tensor]:
    """Explicit CAM backward: returns (dx, dgamma)."""
    b, c, h, w = x.shape
    xm = x.reshape(b, c, -1)
    dym = dy.reshape(b, c, -1)
    e = xm @ xm.transpose(1, 2)
    a = softmax_lastdim(-e)
    o = a @ xm
    dgamma = (dym * o).sum().reshape(1)
    do = gamma * dym
    da = do @ xm.transpose(1, 2)
    de = -a * (da - (a * da).sum(dim=-1, keepdim=True))
    dx = dym + a.transpose(1, 2) @ do + (de + de.transpose(1, 2)) @ xm
    return dx.reshape(b, c, h, w), dgamma


# ----------------------------------------------------------------------------
# generator / discriminator over a state_dict
# ----------------------------------------------------------------------------


def _bn(sd: SD, prefix: str, x: Tensor, training: bool, buffers_out: Optional[SD]) -> Tensor:
    if buffers_out is not None and training:
        buffers_out[prefix + "num_batches_tracked"] = sd[prefix + "num_batches_tracked"] + 1
    return batchnorm(x, sd[prefix + "weight"], sd[prefix + "bias"], sd[prefix + "running_mean"],
                     sd[prefix + "running_var"], training, buffers_out, prefix)


def generator_structure(sd: SD) -> Tuple[int, int, bool]:
    """(num_blocks, layers_per_block, has_attention) inferred from the keys."""
    blocks = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("dense_blocks."))
    layers = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith("dense_blocks.0.layers."))
    return blocks, layers, any(k.startswith("attention_modules.") for k in sd)


def generator_forward(sd: SD, x: Tensor, training: bool = True, buffers_out: Optional[SD] = None,
                      taps: Optional[Dict[str, Tensor]] = None, pam_block: Optional[int] = None) -> Tensor:
    """FlexibleUpsamplingModule.forward, generator.py:230-247.  ``taps`` (if given)
    receives named intermediate activations for per-stage parity checks.  ``pam_block``: evaluate the position attention
    in query blocks of that many rows (:func:`pam_blocked`, forward only) -- the 160x320 grid of BASELINE configs[3]."""
    pam = globals()["pam"] if pam_block is None else (lambda *a: pam_blocked(*a, block=pam_block))
    blocks, layers, has_attn = generator_structure(sd)

    def tap(name, t):
        if taps is not None:
            taps[name] = t
        return t

    x = conv2d(x, sd["initial.0.weight"], None, padding=1)
    x = tap("initial", relu(_bn(sd, "initial.1.", x, training, buffers_out)))
    skips: List[Tensor] = []
    for bi in range(blocks):
        for li in range(layers):
            p = f"dense_blocks.{bi}.layers.{li}."
            y = relu(_bn(sd, p + "bn.", x, training, buffers_out))
            y = conv2d(y, sd[p + "conv.weight"], sd[p + "conv.bias"], padding=1)
            x = torch.cat([x, y], dim=1)
        tap(f"dense{bi}", x)
        if has_attn:
            a = f"attention_modules.{bi}."
            pos = pam(x, sd[a + "position_attention.query.weight"], sd[a + "position_attention.query.bias"],
                      sd[a + "position_attention.key.weight"], sd[a + "position_attention.key.bias"],
                      sd[a + "position_attention.value.weight"], sd[a + "position_attention.value.bias"],
                      sd[a + "position_attention.gamma"])
            ch = cam(x, sd[a + "channel_attention.gamma"])
            tap(f"pam{bi}", pos)
            tap(f"cam{bi}", ch)
            f = conv2d(torch.cat([pos, ch], dim=1), sd[a + "fuse.0.weight"], None, padding=1)
            x = relu(_bn(sd, a + "fuse.1.", f, training, buffers_out))
        skips.append(tap(f"skip{bi}", x))
        if bi != blocks - 1:
            t = f"transition_layers.{bi}.layer."
            y = relu(_bn(sd, t + "0.", x, training, buffers_out))
            x = tap(f"trans{bi}", conv2d(y, sd[t + "2.weight"], sd[t + "2.bias"]))
    x = conv2d(x, sd["upsample.0.weight"], None, padding=1)
    x = bicubic_up2(relu(_bn(sd, "upsample.1.", x, training, buffers_out)))
    tap("up1", x)
    x = conv2d(x, sd["upsample.4.weight"], None, padding=1)
    x = bicubic_up2(relu(_bn(sd, "upsample.5.", x, training, buffers_out)))
    tap("up2", x)
    for i, feat in enumerate(reversed(skips)):
        resized = bilinear_to(feat, (x.shape[2], x.shape[3]))
        x = x + conv2d(resized, sd[f"channel_adjust.{i}.weight"], None)
    tap("fused", x)
    return conv2d(x, sd["final.weight"], sd["final.bias"], padding=1)


def discriminator_forward(sd: SD, x: Tensor) -> Tensor:
    """Discriminator1.forward, discriminator.py:70-77."""
    for i in range(1, 5):
        x = leaky_relu(conv2d(x, sd[f"conv{i}.weight"], sd[f"conv{i}.bias"], stride=2, padding=1))
    x = x.flatten(1)
    x = leaky_relu(x @ sd["fc1.weight"].t() + sd["fc1.bias"])
    return x @ sd["fc2.weight"].t() + sd["fc2.bias"]


# ----------------------------------------------------------------------------
# losses
# ----------------------------------------------------------------------------

# torchvision vgg19 ``features`` layout up to index 20 (relu4_1): conv indices and pools
VGG_CONV_IDX = (0, 2, 5, 7, 10, 12, 14, 16, 19)
VGG_POOL_IDX = (4, 9, 18)


def maxpool2(x: Tensor) -> Tensor:
    b, c, h, w = x.shape
    x = x[:, :, : h // 2 * 2, : w // 2 * 2].reshape(b, c, h // 2, 2, w // 2, 2)
    return x.amax(dim=(3, 5))


def perceptual_loss(vgg_sd: SD, x: Tensor, y: Tensor, feature_layers: Sequence[int] = (1, 6, 11, 20)) -> Tensor:
    """PerceptualLoss.forward, losses.py:63-73: sum of mean-|.| over VGG19 features
    after layer indices ``feature_layers``; inputs repeated to 3 channels, no
    ImageNet normalisation.  ``vgg_sd`` has torchvision keys ``{idx}.weight/bias``."""
    fx = x if x.shape[1] == 3 else x.repeat(1, 3, 1, 1)
    fy = y if y.shape[1] == 3 else y.repeat(1, 3, 1, 1)
    loss = torch.zeros((), dtype=x.dtype)
    wanted = set(feature_layers)
    for idx in range(max(wanted) + 1):
        if idx in VGG_CONV_IDX:
            w, b = vgg_sd[f"{idx}.weight"].to(x.dtype), vgg_sd[f"{idx}.bias"].to(x.dtype)
            fx, fy = conv2d(fx, w, b, padding=1), conv2d(fy, w, b, padding=1)
        elif idx in VGG_POOL_IDX:
            fx, fy = maxpool2(fx), maxpool2(fy)
        else:
            fx, fy = relu(fx), relu(fy)
        if idx in wanted:
            loss = loss + (fx - fy).abs().mean()
    return loss


def tv_loss(x: Tensor, weight: float = 1.0) -> Tensor:
    """TVLoss.forward, losses.py:81-87 (note the extra /batch)."""
    b, c, h, w = x.shape
    h_tv = ((x[:, :, 1:, :] - x[:, :, :-1, :]) ** 2).sum()
    w_tv = ((x[:, :, :, 1:] - x[:, :, :, :-1]) ** 2).sum()
    return weight * 2 * (h_tv / (b * c * (h - 1) * w) + w_tv / (b * c * h * (w - 1))) / b


def gaussian_window(size: int = 11, sigma: float = 1.5, dtype=torch.float32) -> Tensor:
    g = torch.exp(-((torch.arange(size, dtype=torch.float32) - size // 2) ** 2) / (2 * sigma ** 2))
    return (g / g.sum()).to(dtype)


def ssim(img1: Tensor, img2: Tensor, window_size: int = 11) -> Tensor:
    """SSIM.forward (size_average=True), losses.py:98-136: 11x11 Gaussian (sigma 1.5)
    depthwise filtering with zero padding, C1=1e-4, C2=9e-4, mean of the map."""
    c = img1.shape[1]
    g = gaussian_window(window_size, 1.5, torch.float32)
    w2 = (g[:, None] @ g[None, :]).to(img1.dtype)
    win = w2[None, None].expand(c, 1, window_size, window_size).contiguous()
    pad = window_size // 2

    def blur(t):
        return F.conv2d(t, win, padding=pad, groups=c)

    mu1, mu2 = blur(img1), blur(img2)
    s11 = blur(img1 * img1) - mu1 * mu1
    s22 = blur(img2 * img2) - mu2 * mu2
    s12 = blur(img1 * img2) - mu1 * mu2
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu1 * mu2 + c1) * (2 * s12 + c2)) / ((mu1 * mu1 + mu2 * mu2 + c1) * (s11 + s22 + c2))
    return m.mean()


def bce_with_logits(z: Tensor, target: float) -> Tensor:
    """``BCEWithLogitsLoss()`` (mean) against a constant target."""
    return (torch.clamp_min(z, 0) - z * target + torch.log1p(torch.exp(-z.abs()))).mean()


def mse(a: Tensor, b: Tensor) -> Tensor:
    return ((a - b) ** 2).mean()


# ----------------------------------------------------------------------------
# optimiser / schedule
# ----------------------------------------------------------------------------


def adamw_update(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float,
                 beta1: float = 0.5, beta2: float = 0.999, eps: float = 1e-8, wd: float = 1e-4) -> Tuple[Tensor, Tensor, Tensor]:
    """One ``torch.optim.AdamW`` update (``GAN_DANet_train.ipynb:182-183``); step is 1-based."""
    p = p * (1 - lr * wd)
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    denom = (v.sqrt() / math.sqrt(1 - beta2 ** step)) + eps
    p = p - (lr / (1 - beta1 ** step)) * (m / denom)
    return p, m, v


def cosine_warm_restarts_lr(epoch: int, base_lr: float, t0: int = 10, t_mult: int = 2, eta_min: float = 1e-6) -> float:
    """``CosineAnnealingWarmRestarts(T_0=10, T_mult=2, eta_min=1e-6)`` stepped per epoch."""
    t_i, t_cur = t0, epoch
    while t_cur >= t_i:
        t_cur -= t_i
        t_i *= t_mult
    return eta_min + (base_lr - eta_min) * (1 + math.cos(math.pi * t_cur / t_i)) / 2


# ----------------------------------------------------------------------------
# the G+D training step (GAN_DANet_train.ipynb:225-269)
# ----------------------------------------------------------------------------


def prepare_input(lr_grace_05: Tensor, hr_aux: Tensor) -> Tensor:
    """``GAN_DANet_train.ipynb:226-232``: bicubic x0.5 of the 0.5-degree field, bicubic x0.25
    of the aux stack, channel concat -> [B, 1+C_aux, h, w]."""
    return torch.cat([bicubic_down(lr_grace_05, 0.5), bicubic_down(hr_aux, 0.25)], dim=1)


class TrainState:
    """Parameters + AdamW moments of G and D, BN buffers of G, step counter."""

    def __init__(self, g_sd: SD, d_sd: SD, vgg_sd: SD):
        self.g = {k: v.clone() for k, v in g_sd.items()}
        self.d = {k: v.clone() for k, v in d_sd.items()}
        self.vgg = vgg_sd
        self.g_param_names = [k for k in g_sd if not (k.endswith("running_mean") or k.endswith("running_var")
                                                      or k.endswith("num_batches_tracked"))]
        self.d_param_names = list(d_sd.keys())
        self.m_g = {k: torch.zeros_like(self.g[k]) for k in self.g_param_names}
        self.v_g = {k: torch.zeros_like(self.g[k]) for k in self.g_param_names}
        self.m_d = {k: torch.zeros_like(self.d[k]) for k in self.d_param_names}
        self.v_d = {k: torch.zeros_like(self.d[k]) for k in self.d_param_names}
        self.step = 0


def train_step(st: TrainState, lr_grace_05: Tensor, lr_grace_025: Tensor, hr_aux: Tensor, epoch: int, epochs: int,
               lr_d: float = 4e-4, lr_g: float = 2e-4, tv_weight: float = 1e-5,
               grads_out: Optional[Dict[str, SD]] = None) -> Dict[str, float]:
    """One iteration of the hot loop.  Returns the scalar losses; mutates ``st``."""
    st.step += 1
    x_in = prepare_input(lr_grace_05, hr_aux)
    gp = {k: (v.detach().requires_grad_(True) if k in st.m_g else v) for k, v in st.g.items()}
    buffers: SD = {}
    hr = generator_forward(gp, x_in, training=True, buffers_out=buffers)

    # --- discriminator step (:246-256)
    dp = {k: v.detach().requires_grad_(True) for k, v in st.d.items()}
    real_out = discriminator_forward(dp, lr_grace_025)
    fake_out = discriminator_forward(dp, hr.detach())
    loss_d = (bce_with_logits(real_out, 1.0) + bce_with_logits(fake_out, 0.0)) / 2
    d_grads = torch.autograd.grad(loss_d, [dp[k] for k in st.d_param_names])
    for k, g in zip(st.d_param_names, d_grads):
        st.d[k], st.m_d[k], st.v_d[k] = adamw_update(st.d[k], g, st.m_d[k], st.v_d[k], st.step, lr_d)

    # --- generator step (:259-269), discriminator already updated
    fake_out = discriminator_forward(st.d, hr)
    loss_adv = bce_with_logits(fake_out, 1.0)
    loss_pix = mse(hr, lr_grace_025)
    loss_ssim = 1 - ssim(hr.detach(), lr_grace_025)          # evaluated, not in the objective (:263 vs :267)
    loss_tv = tv_loss(hr, tv_weight)
    loss_perc = perceptual_loss(st.vgg, hr, lr_grace_025)
    w = epoch / epochs
    loss_g = (1 - w) * loss_pix + w * loss_adv + loss_tv + loss_perc
    g_grads = torch.autograd.grad(loss_g, [gp[k] for k in st.g_param_names], allow_unused=True)
    if grads_out is not None:
        grads_out["d"] = dict(zip(st.d_param_names, d_grads))
        grads_out["g"] = {k: (g if g is not None else torch.zeros_like(gp[k])) for k, g in zip(st.g_param_names, g_grads)}
        grads_out["hr"] = hr.detach()
    for k, g in zip(st.g_param_names, g_grads):
        if g is None:
            continue
        st.g[k], st.m_g[k], st.v_g[k] = adamw_update(st.g[k], g, st.m_g[k], st.v_g[k], st.step, lr_g)
    for k, v in buffers.items():
        st.g[k] = v
    return {"loss_D": float(loss_d), "loss_G": float(loss_g), "adv": float(loss_adv), "pixel": float(loss_pix),
            "ssim": float(loss_ssim), "tv": float(loss_tv), "perceptual": float(loss_perc)}
