"""CPU oracle (numpy) of the steps AFTER the training path: inference post-processing and deep-ensemble statistics.

TEST INFRASTRUCTURE ONLY: imported by tests/ (and nowhere in the product path).  Each function restates the algorithm of
the reference notebook code it cites; pinned against outputs of the reference's own function text executed in the build
container (oracle/make_golden_post.py -> tests/golden/postprocess.pt, checked in tests/test_oracle_golden.py).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def hist_match(source: np.ndarray, reference: np.ndarray, weight: float = 1.0) -> np.ndarray:
    """mild_histogram_matching (test.ipynb:115-125); weight = 1 is simple_histogram_matching (:104-113).

    Restated on sorted arrays instead of np.unique: the CDF of a value v is #{elements <= v}/n; the reference CDF has one
    knot per distinct reference value; the matched value is the piecewise-linear interpolation (np.interp, float64) of the
    reference knots at the source CDF."""
    shape = source.shape
    s = np.asarray(source).ravel()
    t = np.sort(np.asarray(reference).ravel())
    s_sorted = np.sort(s)
    s_q = np.searchsorted(s_sorted, s, side="right").astype(np.float64) / s.size
    last = np.ones(t.size, dtype=bool)                  # last element of every run of equal reference values
    last[:-1] = t[1:] != t[:-1]
    t_values = t[last]
    t_q = (np.nonzero(last)[0] + 1).astype(np.float64) / t.size
    matched = np.interp(s_q, t_q, t_values)
    return ((1.0 - weight) * s.astype(np.float64) + weight * matched).reshape(shape)


def feather_mask(h: int, w: int, sigma: int = 5) -> np.ndarray:
    """The blend mask of smooth_blend (test.ipynb:487-492): linear ramps of ``sigma`` pixels on the four edges
    (rows first, then columns combined by max), smoothed by a Gaussian of the same sigma."""
    from scipy.ndimage import gaussian_filter
    m = np.ones((h, w), dtype=float)
    up, down = np.linspace(0, 1, sigma), np.linspace(1, 0, sigma)
    m[:sigma, :] = up[:, None]
    m[h - sigma:, :] = down[:, None]
    m[:, :sigma] = np.maximum(m[:, :sigma], up[None, :])
    m[:, w - sigma:] = np.maximum(m[:, w - sigma:], down[None, :])
    return gaussian_filter(m, sigma=sigma)


def smooth_blend(generated: np.ndarray, grace: np.ndarray, region: Tuple[int, int, int, int], sigma: int = 5) -> np.ndarray:
    """smooth_blend (test.ipynb:482-496) on [B, C, H, W] arrays; the mask is cast to float32 as the reference does (:493)."""
    sr, er, sc, ec = region
    m = feather_mask(er - sr, ec - sc, sigma).astype(np.float32)
    out = np.array(generated, dtype=np.float32, copy=True)
    out[:, :, sr:er, sc:ec] = out[:, :, sr:er, sc:ec] * (1 - m) + np.asarray(grace, dtype=np.float32)[:, :, sr:er, sc:ec] * m
    return out


def destandardise(x: np.ndarray, scale: float, mean: float, trend=None, keep=None) -> np.ndarray:
    """(x + trend) -> StandardScaler.inverse_transform (x*scale_ + mean_) -> NaN where the plateau mask is 0
    (test.ipynb:180-191; deep_ensemble.ipynb:415-416)."""
    y = np.asarray(x, dtype=np.float64)
    if trend is not None:
        y = y + trend
    y = y * scale + mean
    if keep is not None:
        y = np.where(np.asarray(keep, dtype=bool), y, np.nan)
    return y


def compute_uncertainty(all_preds: np.ndarray, trues: np.ndarray, keep: np.ndarray):
    """EnsembleTrainer.compute_uncertainty (deep_ensemble.ipynb:430-473).  all_preds [M, T, C, H, W], trues [T, C, H, W],
    keep [H, W] (True inside the plateau, i.e. tpb_h != 0).  Returns (preds_ts [M,T,C], mean_preds [T,C], std_preds [T,C], r2)."""
    keep = np.asarray(keep, dtype=bool)
    p = np.where(keep, np.asarray(all_preds, dtype=np.float64), np.nan)
    t = np.where(keep, np.asarray(trues, dtype=np.float64), np.nan)
    with np.errstate(all="ignore"):
        preds_ts = np.nanmean(p, axis=(3, 4))
        trues_ts = np.nanmean(t, axis=(2, 3))
        mean_preds = np.nanmean(preds_ts, axis=0)
        std_preds = np.nanstd(preds_ts, axis=0)
    ok = ~np.isnan(trues_ts) & ~np.isnan(mean_preds)
    return preds_ts, mean_preds, std_preds, r2_score(trues_ts[ok], mean_preds[ok])


def r2_score(y_true: np.ndarray, y_pred: np.ndarray) -> float:
    """sklearn.metrics.r2_score for one output: 1 - SS_res / SS_tot."""
    y_true, y_pred = np.asarray(y_true, dtype=np.float64).ravel(), np.asarray(y_pred, dtype=np.float64).ravel()
    return float(1.0 - ((y_true - y_pred) ** 2).sum() / ((y_true - y_true.mean()) ** 2).sum())


def pixel_statistics(all_preds: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Per-pixel ensemble mean and spread (np.nanmean / np.nanstd over the member axis, deep_ensemble.ipynb:463-464 applied
    to the fields instead of the spatial means)."""
    with np.errstate(all="ignore"):
        return np.nanmean(np.asarray(all_preds, dtype=np.float64), axis=0), np.nanstd(np.asarray(all_preds, dtype=np.float64), axis=0)
