"""Rows after the training step: inference post-processing (f3) and deep-ensemble statistics (f1, BASELINE.json
configs[4]) on the device, through the C ABI, against the golden outputs of the reference's own notebook functions
(tests/golden/postprocess.pt, made by oracle/make_golden_post.py) and the numpy oracle (oracle/postprocess_oracle.py)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def _require_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gan_danet_b200 import _lib
    _lib.lib_for_device(0)


@pytest.fixture(scope="module")
def post_oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import postprocess_oracle
    return postprocess_oracle


def test_hist_match_golden(golden):
    """apply_mild_histogram_matching / simple_histogram_matching (test.ipynb:104-131): continuous, tied, unequal sizes."""
    from gan_danet_b200 import postprocess as PP
    g = golden("postprocess")
    for case in g["hist"]:
        got = PP.hist_match(case["src"].to(DEV), case["ref"].to(DEV), case["weight"])
        err = float((got.cpu().double() - case["out"]).abs().max())
        assert err < 2e-6 * max(1.0, float(case["out"].abs().max())), (case["weight"], err)       # float32 output of float64 CDF arithmetic
    one = g["hist_simple"]
    got = PP.hist_match(one["src"][None].to(DEV), one["ref"][None].to(DEV), 1.0)[0]
    assert torch.equal(got.cpu(), one["out"].float())                  # weight 1: the output IS a reference value or an exact interpolation


def test_hist_match_full_size(post_oracle):
    """256x512 fields (BASELINE grid), 4 samples; properties: weight 0 is the identity, weight 1 is monotone in the source and
    maps onto the reference's value range; parity with the numpy oracle on every element."""
    from gan_danet_b200 import postprocess as PP
    gen = torch.Generator().manual_seed(3)
    src = torch.randn(4, 1, 256, 512, generator=gen) * 2 + 1
    ref = torch.randn(4, 1, 256, 512, generator=gen).abs()
    s, r = src.to(DEV), ref.to(DEV)
    assert torch.equal(PP.hist_match(s, r, 0.0), s)
    full = PP.hist_match(s, r, 1.0)
    for b in range(4):
        order = torch.argsort(s[b].flatten())
        assert bool((full[b].flatten()[order].diff() >= 0).all())
        assert float(full[b].max()) == float(r[b].max()) and float(full[b].min()) >= float(r[b].min())
    want = np.stack([post_oracle.hist_match(a, b, 0.2) for a, b in zip(src.numpy(), ref.numpy())])
    got = PP.hist_match(s, r, 0.2).cpu().double().numpy()
    assert np.abs(got - want).max() < 2e-6 * np.abs(want).max()


def test_blend_resize_destandardise(golden, post_oracle):
    from gan_danet_b200 import postprocess as PP
    g = golden("postprocess")
    for case in g["blend"]:                                                                  # smooth_blend, test.ipynb:482-496
        a = case["a"].clone().to(DEV)
        out = PP.smooth_blend(a, case["b"].to(DEV), case["region"], case["sigma"])
        assert out.data_ptr() == a.data_ptr()
        assert float((out.cpu() - case["out"]).abs().max()) < 1e-6
    for case in g["resize"]:                                                                 # F.interpolate bicubic, test.ipynb:553,559
        got = PP.bicubic_resize(case["x"].to(DEV), case["scale"])
        assert got.shape == case["out"].shape
        assert rel_err(got, case["out"]) < 5e-6, case["scale"]            # float32 source coordinates and taps (float64 reference)
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(6, 1, 45, 22, generator=gen)
    trend = torch.randn(6, 1, 45, 22, generator=gen)
    keep = torch.rand(45, 22, generator=gen) > 0.3
    got = PP.destandardise(x.to(DEV), 7.25, -1.5, trend.to(DEV), keep.to(DEV)).cpu().numpy()  # test.ipynb:180-191
    want = post_oracle.destandardise(x.numpy(), 7.25, -1.5, trend.numpy(), keep.numpy())
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.nanmax(np.abs(got - want)) < 1e-5


def test_uncertainty_golden(golden, post_oracle):
    """EnsembleTrainer.compute_uncertainty (deep_ensemble.ipynb:430-473) against the reference's own output, NaN inside the
    kept region included; per-pixel statistics against np.nanmean / np.nanstd."""
    from gan_danet_b200.ensemble import EnsembleTrainer
    import tempfile
    u = golden("postprocess")["uncertainty"]
    ens = EnsembleTrainer(u["preds"].shape[0], {}, ensemble_dir=tempfile.mkdtemp())
    mean_preds, std_preds, r2 = ens.compute_uncertainty(u["preds"].to(DEV), u["trues"].to(DEV), u["keep"].to(DEV))
    assert np.abs(mean_preds - u["mean_preds"].numpy()).max() < 2e-6
    assert np.abs(std_preds - u["std_preds"].numpy()).max() < 2e-6
    assert abs(r2 - u["r2"]) < 1e-5
    pm, ps = ens.pixel_statistics(u["preds"].to(DEV))
    wm, ws = post_oracle.pixel_statistics(u["preds"].numpy())
    assert np.array_equal(np.isnan(pm.cpu().numpy()), np.isnan(wm))
    assert np.nanmax(np.abs(pm.cpu().numpy() - wm)) < 1e-6 and np.nanmax(np.abs(ps.cpu().numpy() - ws)) < 1e-6
    # all-masked rows -> NaN like np.nanmean; more than 16 members takes the re-reading kernel
    from gan_danet_b200 import postprocess as PP
    assert torch.isnan(PP.masked_spatial_mean(u["trues"].to(DEV), torch.zeros_like(u["keep"]).to(DEV))).all()
    big = torch.randn(20, 3, 1000, generator=torch.Generator().manual_seed(1))
    m20, s20 = PP.ensemble_stats(big.to(DEV), 2.0, 0.5)
    assert rel_err(m20, (big.double() * 2 + 0.5).mean(0)) < 1e-6 and rel_err(s20, (big.double() * 2 + 0.5).std(0, unbiased=False)) < 1e-6


def test_ensemble_sweep_vs_oracle(oracle, post_oracle, tmp_path):
    """cfg5 in small: 3 members (seeds 42..44), T = 6 months of the authors' grid family (C_in 46, 8x16 -> 32x64), eval-mode sweep
    + inverse scaling + masked means + member statistics on the device against the CPU oracle generator and numpy."""
    import gan_danet_b200 as P
    from gan_danet_b200 import engine as E
    from gan_danet_b200.ensemble import EnsembleTrainer
    from gan_danet_b200.synthetic import make_batch
    M, T, h, w = 3, 6, 8, 16
    ens = EnsembleTrainer(M, {}, ensemble_dir=str(tmp_path))
    assert ens.seeds == [42, 43, 44] and ens.local_members() == [0, 1, 2]
    for i in range(M):                                                          # members as train_ensemble would leave them on disk
        ens.set_seed(ens.seeds[i])
        G = P.FlexibleUpsamplingModule(46, attention_type="danet")
        G.apply(P.weights_init_normal)
        with torch.no_grad():
            for n, p in G.named_parameters():
                if n.endswith("gamma"):
                    p.fill_(0.05)
            for mod in G.modules():
                if isinstance(mod, torch.nn.BatchNorm2d):
                    mod.running_mean.normal_(0.0, 0.1)
                    mod.running_var.uniform_(0.5, 1.5)
        torch.save(G.state_dict(), ens.member_path(i))
    old = E.conv_precision
    E.set_conv_precision("fp32")
    try:
        models = ens.load_ensemble_models(P.FlexibleUpsamplingModule, torch.device(DEV), 46, "danet")
        for m_ in models:
            m_.set_pam_precision("fp32")
        lr05, lr025, aux = make_batch(0, T, h, w)
        loader = [(lr05[:4], lr025[:4], aux[:4]), (lr05[4:], lr025[4:], aux[4:])]
        scale, mean = 9.5, 1.25
        all_preds, trues = ens.predict_ensemble(models, loader, scaler=(scale, mean))
    finally:
        E.set_conv_precision(old)
    assert all_preds.shape == (M, T, 1, 4 * h, 4 * w) and trues.shape == (T, 1, 4 * h, 4 * w)
    x = oracle.prepare_input(lr05, aux)
    want = []
    for i in range(M):
        sd = torch.load(ens.member_path(i), map_location="cpu")
        want.append(oracle.generator_forward({k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}, x.double(), training=False) * scale + mean)
    want = torch.stack(want)
    assert rel_err(all_preds, want) < 1e-4, rel_err(all_preds, want)
    assert rel_err(trues, lr025.double() * scale + mean) < 1e-6
    keep = torch.rand(4 * h, 4 * w, generator=torch.Generator().manual_seed(9)) > 0.4
    mean_preds, std_preds, r2 = ens.compute_uncertainty(all_preds, trues, keep.to(DEV))
    _, wm, wsd, wr2 = post_oracle.compute_uncertainty(want.numpy(), (lr025.double() * scale + mean).numpy(), keep.numpy())
    assert np.abs(mean_preds - wm).max() < 1e-3 * np.abs(wm).max() and np.abs(std_preds - wsd).max() < 2e-3 * np.abs(wsd).max() + 1e-5
    assert abs(r2 - wr2) < 1e-3 * max(1.0, abs(wr2))      # the standardised targets have nearly constant spatial means: r2 is large and negative


@pytest.mark.parametrize("hw,conv", [((8, 16), "fp32"), ((64, 128), "bf16")])
def test_graphed_generator_matches_eager(hw, conv):
    """the eval forward captured as one CUDA graph replays bit-identically to the eager launches, for new inputs
    copied into the static buffers and after an in-place weight update."""
    import gan_danet_b200 as P
    from gan_danet_b200 import engine as E
    from gan_danet_b200.inference import GraphedGenerator
    from gan_danet_b200.synthetic import fast_batch
    from gan_danet_b200.trainer import generator_forward_nhwc, prepare_input_nhwc
    torch.manual_seed(0)
    G = P.FlexibleUpsamplingModule(46)
    G.apply(P.weights_init_normal)
    with torch.no_grad():
        for n, p in G.named_parameters():
            if n.endswith("gamma"):
                p.fill_(0.05)
    G = G.to(DEV).eval()
    old = E.conv_precision
    E.set_conv_precision(conv)
    try:
        a = [t.to(DEV) for t in fast_batch(1, 2, *hw)]
        b = [t.to(DEV) for t in fast_batch(2, 2, *hw)]
        gg = GraphedGenerator(G, a[0], a[2])
        assert gg.launches_captured > 100

        def eager(lr05, aux):
            with torch.no_grad():
                return generator_forward_nhwc(G, prepare_input_nhwc(lr05, aux))

        assert torch.equal(gg(a[0], a[2]), eager(a[0], a[2]))
        assert torch.equal(gg(b[0], b[2]), eager(b[0], b[2]))
        with torch.no_grad():
            G.final.weight.mul_(1.5)
        yb = gg(b[0], b[2])
        assert torch.equal(yb, eager(b[0], b[2])) and torch.isfinite(yb).all()
        with pytest.raises(Exception):
            gg(a[0][:1], a[2][:1])
        G.train()
        with pytest.raises(Exception):
            GraphedGenerator(G, a[0], a[2])
    finally:
        E.set_conv_precision(old)


def test_train_ensemble_small(tmp_path):
    """EnsembleTrainer.train_ensemble (deep_ensemble.ipynb:322-340) in small: two members (seeds 42, 43) trained for two epochs of two batches
    on this rank, checkpoints written under the notebook's file names, members differ, and they load back through load_ensemble_models."""
    import gan_danet_b200 as P
    from gan_danet_b200.ensemble import EnsembleTrainer
    from gan_danet_b200.synthetic import make_batch
    ens = EnsembleTrainer(2, dict(input_channels=46, attention_type="danet", epochs=2, perceptual=False, sample_hw=(32, 64)), ensemble_dir=str(tmp_path))

    def batches(epoch):
        return [tuple(t.to(DEV) for t in make_batch(4 * epoch + 2 * i, 2, 8, 16)) for i in range(2)]

    curves = ens.train_ensemble(batches, epochs=2)
    assert set(curves) == {0, 1} and all(len(c) == 4 and all(np.isfinite(c)) for c in curves.values())
    assert curves[0] != curves[1]                                   # different seeds, same data split
    assert os.path.basename(ens.member_path(0)) == "best_model_member_1.pth" and os.path.exists(ens.member_path(1))
    models = ens.load_ensemble_models(P.FlexibleUpsamplingModule, torch.device(DEV), 46, "danet")
    assert len(models) == 2 and not models[0].training
    assert not torch.equal(models[0].state_dict()["final.weight"], models[1].state_dict()["final.weight"])


@pytest.mark.parametrize("rows,n", [(3, 1), (2, 5), (4, 1000), (5, 8192), (3, 15840), (2, 131072), (181, 20000)])
def test_sort_rows(rows, n):
    """gdn_sort_rows (the np.sort / np.unique step of test.ipynb:117-121, own bitonic kernel) against torch.sort: bitwise equal values, NaNs last."""
    from gan_danet_b200 import postprocess as PP
    g = torch.Generator().manual_seed(rows * 7 + n)
    x = torch.randn(rows, n, generator=g)
    x[0, 0] = float("nan")
    if n > 4:
        x[-1, 1] = float("nan"); x[0, 2] = -0.0; x[0, 3] = 0.0; x[-1, 4] = float("inf"); x[0, n // 2] = -float("inf")
        x[-1, : n // 3] = x[-1, n // 3]          # a long run of equal values
    x = x.to(DEV)
    got = PP.sort_rows(x)
    want = torch.sort(x, dim=1).values
    torch.cuda.synchronize()
    assert torch.equal(torch.isnan(got), torch.isnan(want))
    assert torch.equal(torch.nan_to_num(got, nan=0.0), torch.nan_to_num(want, nan=0.0))      # -0.0 == 0.0 compare equal, as in np.sort's output order
