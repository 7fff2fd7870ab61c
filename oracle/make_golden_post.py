"""Golden fixtures for the rows AFTER the training step (f1 ensemble statistics, f3 inference post-processing).

The reference keeps these functions inside notebooks, so they cannot be imported: this script reads the notebook JSON
(read-only, /root/reference), cuts the UNMODIFIED source text of each function / method out of its code cell, ``exec``s that
text, and runs it on seeded synthetic inputs.  File access inside the executed text (``np.load('tpb_h.npy')``) is served
from the synthetic mask.  Run in the build container only:

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_post.py

Writes tests/golden/postprocess.pt (pins oracle/postprocess_oracle.py in tests/test_oracle_golden.py).
"""
from __future__ import annotations

import json
import os
import re
import sys
import textwrap

os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402


def cell_text(notebook: str, needle: str) -> str:
    nb = json.load(open(os.path.join("/root/reference", notebook)))
    for c in nb["cells"]:
        src = "".join(c["source"])
        if c["cell_type"] == "code" and needle in src:
            return src
    raise KeyError(needle)


def cut_def(text: str, name: str, indent: int = 0) -> str:
    """Source of ``def name`` at the given indentation, up to the next statement at the same or a lower indentation."""
    pad = " " * indent
    m = re.search(rf"^{pad}def {name}\(", text, re.M)
    lines = text[m.start():].split("\n")
    out = [lines[0]]
    for ln in lines[1:]:
        if ln.strip() and (len(ln) - len(ln.lstrip())) <= indent:
            break
        out.append(ln)
    return textwrap.dedent("\n".join(out))


def main():
    from scipy.ndimage import gaussian_filter
    from sklearn.metrics import r2_score
    g = torch.Generator().manual_seed(2024)
    ns = {"np": np, "torch": torch, "F": F, "gaussian_filter": gaussian_filter, "r2_score": r2_score, "device": torch.device("cpu")}
    t1 = cell_text("test.ipynb", "def mild_histogram_matching")
    for fn in ("simple_histogram_matching", "mild_histogram_matching", "apply_mild_histogram_matching"):
        exec(cut_def(t1, fn), ns)
    exec(cut_def(cell_text("test.ipynb", "def smooth_blend"), "smooth_blend"), ns)
    exec(cut_def(cell_text("deep_ensemble.ipynb", "class EnsembleTrainer"), "compute_uncertainty", indent=4), ns)

    gold = {}
    # --- histogram matching (test.ipynb:104-131): continuous values, and heavily tied values (quantised fields)
    src = torch.randn(3, 1, 24, 20, generator=g) * 1.3 + 0.2
    ref = torch.randn(3, 1, 24, 20, generator=g) * 0.7 - 0.1
    src_q = torch.round(src * 4) / 4
    ref_q = torch.round(ref * 3) / 3
    ref_small = torch.randn(3, 1, 12, 10, generator=g)
    gold["hist"] = []
    for s, r, w in ((src, ref, 0.2), (src, ref, 1.0), (src_q, ref_q, 0.2), (src_q, ref, 1.0), (src, ref_q, 0.5), (src, ref_small, 0.2), (src, ref, 0.0)):
        out = ns["apply_mild_histogram_matching"](s, r, weight=w)
        gold["hist"].append({"src": s, "ref": r, "weight": w, "out": out.double()})
    one = ns["simple_histogram_matching"](src[0].numpy(), ref[0].numpy())
    gold["hist_simple"] = {"src": src[0], "ref": ref[0], "out": torch.from_numpy(np.asarray(one, dtype=np.float64))}

    # --- smooth_blend (test.ipynb:482-496)
    a = torch.randn(2, 1, 40, 56, generator=g)
    b = torch.randn(2, 1, 40, 56, generator=g)
    gold["blend"] = []
    for region, sigma in (((4, 30, 6, 50), 5), ((0, 40, 0, 56), 5), ((10, 25, 20, 41), 3)):
        out = ns["smooth_blend"](a.clone(), b, region=region, sigma=sigma)
        gold["blend"].append({"a": a, "b": b, "region": region, "sigma": sigma, "out": out})

    # --- bicubic x1.25 and x4 of the inference pipeline (test.ipynb:553,559): ATen itself
    y = torch.randn(2, 1, 22, 45, generator=g)
    gold["resize"] = [{"x": y, "scale": sf, "out": F.interpolate(y.double(), scale_factor=sf, mode="bicubic", align_corners=False)} for sf in (1.25, 4, 0.5)]

    # --- EnsembleTrainer.compute_uncertainty (deep_ensemble.ipynb:430-473): M members, T months, 1 channel
    M, T, H, W = 5, 7, 18, 11
    keep = (torch.rand(H, W, generator=g) > 0.35)
    tpb = keep.to(torch.float32).numpy()                              # tpb_h.npy: 0 = outside the plateau
    ns["np"] = type("NPProxy", (), {"__getattr__": lambda self, k: getattr(np, k), "load": staticmethod(lambda path: tpb)})()
    trues = (torch.randn(T, 1, H, W, generator=g) * 3.0 + 1.0).numpy()
    preds = np.stack([trues + 0.4 * torch.randn(T, 1, H, W, generator=g).numpy() + 0.1 * m for m in range(M)], 0).astype(np.float32)
    preds[1, 2, 0, 3, 4] = np.nan                                      # a NaN inside the kept region must be skipped (np.nanmean)
    mean_preds, std_preds, r2 = ns["compute_uncertainty"](None, preds, trues)
    gold["uncertainty"] = {"preds": torch.from_numpy(preds), "trues": torch.from_numpy(trues), "keep": keep,
                           "mean_preds": torch.from_numpy(mean_preds).double(), "std_preds": torch.from_numpy(std_preds).double(), "r2": float(r2)}
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, "postprocess.pt")
    torch.save(gold, path)
    print(f"postprocess: {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
