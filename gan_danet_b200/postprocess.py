"""Inference post-processing and ensemble statistics on the device (f1, f3) -- thin callers of the C ABI.

Replaces the CPU numpy code that follows the generator in the reference's product pipeline:
``test.ipynb:104-131`` (histogram matching), ``:175-191`` (+trend, inverse standardisation, plateau mask -> NaN),
``:482-496`` (feathered region blend), ``:553,559`` (bicubic x1.25 / x4) and ``deep_ensemble.ipynb:415-416,450-467``
(inverse scaling, masked spatial means, mean / std over the members).  Every function takes CUDA float32 tensors and
raises on anything else: there is no CPU path (the numpy restatement lives in ``oracle/postprocess_oracle.py`` and is
test infrastructure).  ``torch.sort`` provides the ascending copies the histogram kernel searches (plumbing).
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import torch

from . import _lib as L
from . import engine as E


def _f32(t: torch.Tensor, what: str) -> torch.Tensor:
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise L.GdnError(f"{what}: expected a CUDA tensor (there is no CPU path)")
    return t.to(torch.float32).contiguous()


def _keep_bytes(keep: Optional[torch.Tensor], like: torch.Tensor, hw: int) -> Optional[torch.Tensor]:
    if keep is None:
        return None
    k = torch.as_tensor(keep).to(device=like.device)
    k = (k != 0).to(torch.uint8).contiguous()
    if k.numel() != hw:
        raise L.GdnError(f"mask has {k.numel()} pixels, fields have {hw}")
    return k


def destandardise(x: torch.Tensor, scale: float, mean: float, trend: Optional[torch.Tensor] = None, keep: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``scaler.inverse_transform(x + trend)`` with ``NaN`` where ``keep == 0``; x [..., H, W], trend broadcastable to x
    (same shape after expansion), keep [H, W].  test.ipynb:180-191, deep_ensemble.ipynb:415-416."""
    x = _f32(x, "destandardise")
    hw = x.shape[-1] * x.shape[-2]
    t = _f32(trend.expand_as(x), "destandardise(trend)") if trend is not None else None
    k = _keep_bytes(keep, x, hw)
    out = torch.empty_like(x)
    L.check(E._lib(x).gdn_destandardise(x.data_ptr(), t.data_ptr() if t is not None else None, k.data_ptr() if k is not None else None, out.data_ptr(),
                                        x.numel() // hw, hw, float(scale), float(mean), E._stream()), "gdn_destandardise")
    return out


def masked_spatial_mean(x: torch.Tensor, keep: Optional[torch.Tensor] = None, scale: float = 1.0, shift: float = 0.0) -> torch.Tensor:
    """``np.nanmean(where(keep, x*scale+shift, nan), axis=(-2, -1))``: x [..., H, W] -> [...] (deep_ensemble.ipynb:450-460)."""
    x = _f32(x, "masked_spatial_mean")
    hw = x.shape[-1] * x.shape[-2]
    k = _keep_bytes(keep, x, hw)
    out = torch.empty(x.shape[:-2], dtype=torch.float32, device=x.device)
    lib, rows = E._lib(x), x.numel() // hw
    ws = E.workspace("spatial_mean", lib.gdn_masked_spatial_mean_ws_bytes(rows, hw), x.device)
    L.check(lib.gdn_masked_spatial_mean(x.data_ptr(), k.data_ptr() if k is not None else None, rows, hw, float(scale), float(shift),
                                        out.data_ptr(), ws.data_ptr(), ws.numel(), E._stream()), "gdn_masked_spatial_mean")
    return out


def ensemble_stats(preds: torch.Tensor, scale: float = 1.0, shift: float = 0.0, want_std: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """``np.nanmean`` / ``np.nanstd`` (ddof 0) over the leading member axis of preds [M, ...] (deep_ensemble.ipynb:463-464)."""
    preds = _f32(preds, "ensemble_stats")
    M = preds.shape[0]
    n = preds.numel() // M
    mean = torch.empty(preds.shape[1:], dtype=torch.float32, device=preds.device)
    std = torch.empty_like(mean) if want_std else None
    L.check(E._lib(preds).gdn_ensemble_stats(preds.data_ptr(), n, M, n, float(scale), float(shift), mean.data_ptr(), std.data_ptr() if want_std else None,
                                             E._stream()), "gdn_ensemble_stats")
    return mean, std


def sort_rows(x: torch.Tensor) -> torch.Tensor:
    """Ascending sort of every row of a [rows, n] float32 tensor (NaNs last, as np.sort): ``gdn_sort_rows``, one CTA per row."""
    x = _f32(x, "sort_rows")
    rows, n = x.shape
    lib = E._lib(x)
    out = torch.empty_like(x)
    ws = E.workspace("sort", lib.gdn_sort_rows_ws_bytes(rows, n), x.device)
    L.check(lib.gdn_sort_rows(x.data_ptr(), out.data_ptr(), rows, n, ws.data_ptr(), ws.numel(), E._stream()), "gdn_sort_rows")
    return out


def hist_match(source: torch.Tensor, reference: torch.Tensor, weight: float = 0.2) -> torch.Tensor:
    """``apply_mild_histogram_matching`` (test.ipynb:115-131): per sample (leading axis) the source values are moved a
    fraction ``weight`` of the way to the reference distribution; ``weight = 1`` is ``simple_histogram_matching``."""
    src = _f32(source, "hist_match")
    ref = _f32(reference, "hist_match(reference)")
    B = src.shape[0]
    if ref.shape[0] != B:
        raise L.GdnError("hist_match: source and reference need the same number of samples")
    s2, r2 = src.reshape(B, -1), ref.reshape(B, -1)
    s_sorted, r_sorted = sort_rows(s2), sort_rows(r2)
    out = torch.empty_like(s2)
    L.check(E._lib(src).gdn_hist_match(s2.data_ptr(), s_sorted.data_ptr(), r_sorted.data_ptr(), out.data_ptr(), B, s2.shape[1], r2.shape[1], float(weight),
                                       E._stream()), "gdn_hist_match")
    return out.reshape(src.shape)


def bicubic_resize(x: torch.Tensor, scale_factor: float) -> torch.Tensor:
    """``F.interpolate(x, scale_factor=sf, mode='bicubic', align_corners=False)`` on [B, C, H, W] (test.ipynb:553,559)."""
    x = _f32(x, "bicubic_resize")
    B, Cc, H, W = x.shape
    Ho, Wo = int(math.floor(H * scale_factor)), int(math.floor(W * scale_factor))
    y = torch.empty((B, Cc, Ho, Wo), dtype=torch.float32, device=x.device)
    L.check(E._lib(x).gdn_bicubic_resize(x.data_ptr(), y.data_ptr(), B * Cc, H, W, Ho, Wo, float(scale_factor), float(scale_factor), E._stream()), "gdn_bicubic_resize")
    return y


def feather_mask(h: int, w: int, sigma: int = 5):
    """Host constant of ``smooth_blend`` (test.ipynb:487-492), float32 [h, w]: edge ramps smoothed by scipy's gaussian_filter
    exactly as the reference builds it (computed once per region, not per batch)."""
    import numpy as np
    from scipy.ndimage import gaussian_filter
    m = np.ones((h, w), dtype=float)
    m[0:sigma, :] = np.linspace(0, 1, sigma)[:, None]
    m[-sigma:, :] = np.linspace(1, 0, sigma)[:, None]
    m[:, 0:sigma] = np.maximum(m[:, 0:sigma], np.linspace(0, 1, sigma)[None, :])
    m[:, -sigma:] = np.maximum(m[:, -sigma:], np.linspace(1, 0, sigma)[None, :])
    return gaussian_filter(m, sigma=sigma).astype(np.float32)


_mask_cache = {}


def smooth_blend(hr_generated: torch.Tensor, hr_grace: torch.Tensor, region: Sequence[int], sigma: int = 5) -> torch.Tensor:
    """``smooth_blend`` (test.ipynb:482-496): feathers ``hr_grace`` into the rectangle (sr, er, sc, ec) of ``hr_generated``
    IN PLACE (as the reference assigns into its argument) and returns it."""
    if not (hr_generated.is_cuda and hr_generated.dtype == torch.float32 and hr_generated.is_contiguous()):
        raise L.GdnError("smooth_blend: hr_generated must be a contiguous CUDA float32 tensor (it is updated in place)")
    b = _f32(hr_grace, "smooth_blend(hr_grace)")
    sr, er, sc, ec = (int(v) for v in region)
    key = (er - sr, ec - sc, sigma, hr_generated.device)
    if key not in _mask_cache:
        _mask_cache[key] = torch.from_numpy(feather_mask(er - sr, ec - sc, sigma)).to(hr_generated.device)
    m = _mask_cache[key]
    H, W = hr_generated.shape[-2:]
    L.check(E._lib(b).gdn_blend_region(hr_generated.data_ptr(), b.data_ptr(), m.data_ptr(), hr_generated.data_ptr(), hr_generated.numel() // (H * W), H, W,
                                       sr, sc, er - sr, ec - sc, E._stream()), "gdn_blend_region")
    return hr_generated
