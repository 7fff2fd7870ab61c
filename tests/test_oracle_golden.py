"""Pins the CPU oracle (oracle/gan_danet_oracle.py) against golden vectors produced by the reference itself
(oracle/make_golden.py, run in the build container where /root/reference exists).  CPU only, float64."""
import os
import sys

import pytest
import torch

from conftest import rel_err

TOL = 2e-6   # both sides are float64 computations stored as float32


def _d(sd):
    return {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}


def _grads(out, wrt, r):
    return torch.autograd.grad((out * r.double()).sum(), wrt, allow_unused=True)


@pytest.mark.parametrize("name", ["pam_c160_8x16", "pam_c184_4x8"])
def test_pam(golden, oracle, name):
    g = golden(name)
    sd = {k: v.double().requires_grad_(True) for k, v in g["sd"].items()}
    x = g["x"].double().requires_grad_(True)
    y = oracle.pam(x, sd["query.weight"], sd["query.bias"], sd["key.weight"], sd["key.bias"], sd["value.weight"], sd["value.bias"], sd["gamma"])
    assert rel_err(y, g["y"]) < TOL
    yb = oracle.pam_blocked(x, sd["query.weight"], sd["query.bias"], sd["key.weight"], sd["key.bias"], sd["value.weight"], sd["value.bias"], sd["gamma"], block=32)
    assert rel_err(yb, g["y"]) < TOL
    names = [k for k in g["grads"]]
    grads = _grads(y, [x] + [sd[k] for k in names], g["r"])
    assert rel_err(grads[0], g["dx"]) < TOL
    for k, gr in zip(names, grads[1:]):
        if k == "key.bias":      # analytically zero (SURVEY appendix A identity 3): compare absolutely
            assert gr.abs().max() < 1e-9 * g["grads"]["key.weight"].abs().max().clamp_min(1.0) + 1e-9
        else:
            assert rel_err(gr, g["grads"][k]) < 5e-6, k


@pytest.mark.parametrize("name", ["pam_c160_8x16", "pam_c184_4x8"])
def test_pam_rows(golden, oracle, name):
    """Row-restricted restatement used as the oracle of the 160x320 grid (BASELINE configs[3]): its rows equal the
    reference's output rows, and its gradients equal the reference's under a cotangent that is zero on the other rows."""
    g = golden(name)
    sd = {k: v.double().requires_grad_(True) for k, v in g["sd"].items()}
    x = g["x"].double().requires_grad_(True)
    args = [sd["query.weight"], sd["query.bias"], sd["key.weight"], sd["key.bias"], sd["value.weight"], sd["value.bias"], sd["gamma"]]
    b, c, h, w = x.shape
    rows = torch.tensor([0, 3, h * w // 2, h * w - 1])
    yr = oracle.pam_rows(x, *args, rows)
    assert rel_err(yr, g["y"].reshape(b, c, -1)[:, :, rows]) < TOL
    r = torch.zeros(b, c, h * w, dtype=torch.float64)
    r[:, :, rows] = g["r"].double().reshape(b, c, -1)[:, :, rows]
    full = oracle.pam(x, *args)
    want = torch.autograd.grad((full.reshape(b, c, -1) * r).sum(), [x] + args[:-1])
    got = torch.autograd.grad((yr * r[:, :, rows]).sum(), [x] + args[:-1])
    for a_, b_ in zip(got, want):
        assert (a_ - b_).abs().max() < 1e-9 * b_.abs().max().clamp_min(1.0)


@pytest.mark.parametrize("name", ["pam_c160_8x16"])
def test_pam_flash_restatement(golden, oracle, name):
    """The online-softmax forward/backward restatement the CUDA kernels implement."""
    g = golden(name)
    sd = _d(g["sd"])
    x = g["x"].double()
    b, c, h, w = x.shape
    n = h * w
    q = oracle.conv2d(x, sd["query.weight"], sd["query.bias"]).reshape(b, -1, n).transpose(1, 2)
    k = oracle.conv2d(x, sd["key.weight"], sd["key.bias"]).reshape(b, -1, n).transpose(1, 2)
    v = oracle.conv2d(x, sd["value.weight"], sd["value.bias"]).reshape(b, -1, n).transpose(1, 2)
    o, lse = oracle.pam_core_flash(q, k, v, block=32)
    y = (sd["gamma"] * o).transpose(1, 2).reshape(b, c, h, w) + x
    assert rel_err(y, g["y"]) < TOL
    qa, ka, va = (t.clone().requires_grad_(True) for t in (q, k, v))
    s = torch.softmax(qa @ ka.transpose(1, 2), dim=-1) @ va
    do = torch.randn(o.shape, generator=torch.Generator().manual_seed(0), dtype=torch.float64)
    rq, rk, rv = torch.autograd.grad((s * do).sum(), [qa, ka, va])
    dq, dk, dv = oracle.pam_core_backward(q, k, v, o, lse, do, block=32)
    assert rel_err(dq, rq) < 1e-9 and rel_err(dk, rk) < 1e-9 and rel_err(dv, rv) < 1e-9


@pytest.mark.parametrize("name", ["cam_c160_8x16", "cam_c184_4x8"])
def test_cam(golden, oracle, name):
    g = golden(name)
    x = g["x"].double().requires_grad_(True)
    gamma = g["sd"]["gamma"].double().requires_grad_(True)
    y = oracle.cam(x, gamma)
    assert rel_err(y, g["y"]) < TOL
    dx, dg = _grads(y, [x, gamma], g["r"])
    assert rel_err(dx, g["dx"]) < 5e-6
    assert rel_err(dg, g["grads"]["gamma"]) < 5e-6
    dx2, dg2 = oracle.cam_backward(x.detach(), gamma.detach(), g["r"].double())
    assert rel_err(dx2, g["dx"]) < 5e-6 and rel_err(dg2, g["grads"]["gamma"]) < 5e-6


def _generator_sd(seed, gamma, cin=46):
    import gan_danet_b200 as P
    from gan_danet_b200.models.generator import CAMModule, PAMModule
    torch.manual_seed(seed)
    G = P.FlexibleUpsamplingModule(cin)
    G.apply(P.weights_init_normal)
    with torch.no_grad():
        for m in G.modules():
            if isinstance(m, (PAMModule, CAMModule)):
                m.gamma.fill_(gamma)
    return G


def test_generator(golden, oracle):
    g = golden("generator_cin46_8x16")
    G = _generator_sd(g["seed"], g["gamma"])
    assert list(G.state_dict().keys()) == g["keys"]
    assert sum(p.numel() for p in G.parameters()) == g["n_params"] == 2271993
    names = [k for k, _ in G.named_parameters()]
    sd = {k: (v.double().requires_grad_(True) if k in names else v.double() if v.is_floating_point() else v.clone()) for k, v in G.state_dict().items()}
    x = g["x"].double().requires_grad_(True)
    bufs = {}
    y = oracle.generator_forward(sd, x, training=True, buffers_out=bufs)
    assert rel_err(y, g["y"]) < TOL
    grads = _grads(y, [x] + [sd[k] for k in names], g["r"])
    assert rel_err(grads[0], g["dx"]) < 1e-5
    for k, gr in zip(names, grads[1:]):
        ref_norm = g["grad_norms"][k]
        if "key.bias" in k:
            continue
        assert abs(float(gr.norm()) - ref_norm) <= 1e-5 * max(ref_norm, 1e-12) + 1e-12, k
        if k in g["grads_small"]:
            assert rel_err(gr, g["grads_small"][k]) < 1e-4, k
    for k, v in g["buffers_after"].items():
        assert rel_err(bufs[k], v) < TOL, k


def test_danet_dense_transition(golden, oracle):
    import gan_danet_b200 as P
    from gan_danet_b200.models import generator as PG
    # dense block
    g = golden("denseblock_64_8x16")
    sd = {("dense_blocks.0." + k): v.double() for k, v in g["sd"].items()}
    x = g["x"].double().requires_grad_(True)
    cur = x
    for li in range(4):
        p = f"dense_blocks.0.layers.{li}."
        yb = oracle.relu(oracle.batchnorm(cur, sd[p + "bn.weight"], sd[p + "bn.bias"], sd[p + "bn.running_mean"], sd[p + "bn.running_var"], True))
        cur = torch.cat([cur, oracle.conv2d(yb, sd[p + "conv.weight"], sd[p + "conv.bias"], padding=1)], dim=1)
    assert rel_err(cur, g["y"]) < TOL
    (dx,) = _grads(cur, [x], g["r"])
    assert rel_err(dx, g["dx"]) < 1e-5
    # transition
    g = golden("transition_160_8x16")
    sd = _d(g["sd"])
    x = g["x"].double().requires_grad_(True)
    yb = oracle.relu(oracle.batchnorm(x, sd["layer.0.weight"], sd["layer.0.bias"], sd["layer.0.running_mean"], sd["layer.0.running_var"], True))
    y = oracle.conv2d(yb, sd["layer.2.weight"], sd["layer.2.bias"])
    assert rel_err(y, g["y"]) < TOL
    # DANet block: fuse.0.weight is regenerated from the seed through the mirror package's identical initialisation
    g = golden("danet_c160_8x16")
    torch.manual_seed(0)
    da = PG.DANetAttention(160)
    da.apply(P.weights_init_normal)
    sd = _d(g["sd"])
    sd["fuse.0.weight"] = da.fuse[0].weight.detach().double()
    assert torch.equal(da.position_attention.query.weight.detach(), g["sd"]["position_attention.query.weight"])
    x = g["x"].double().requires_grad_(True)
    pa = "position_attention."
    pos = oracle.pam(x, sd[pa + "query.weight"], sd[pa + "query.bias"], sd[pa + "key.weight"], sd[pa + "key.bias"], sd[pa + "value.weight"],
                     sd[pa + "value.bias"], sd[pa + "gamma"])
    ch = oracle.cam(x, sd["channel_attention.gamma"])
    f = oracle.conv2d(torch.cat([pos, ch], 1), sd["fuse.0.weight"], None, padding=1)
    y = oracle.relu(oracle.batchnorm(f, sd["fuse.1.weight"], sd["fuse.1.bias"], sd["fuse.1.running_mean"], sd["fuse.1.running_var"], True))
    assert rel_err(y, g["y"]) < TOL
    (dx,) = _grads(y, [x], g["r"])
    assert rel_err(dx, g["dx"]) < 1e-5


def _discriminator(seed, sample):
    import gan_danet_b200 as P
    torch.manual_seed(seed)
    D = P.Discriminator1()
    for mod in (D.conv1, D.conv2, D.conv3, D.conv4, D.fc2):
        mod.apply(P.weights_init_normal)
    D._materialise_fc1(sample)
    return D


def test_discriminator(golden, oracle):
    g = golden("discriminator_64x128")
    D = _discriminator(g["seed"], g["x"])
    assert list(D.state_dict().keys()) == g["keys"]
    sd = {k: v.double().requires_grad_(True) for k, v in D.state_dict().items()}
    x = g["x"].double().requires_grad_(True)
    z = oracle.discriminator_forward(sd, x)
    assert rel_err(z, g["logits"]) < TOL
    names = list(sd.keys())
    grads = torch.autograd.grad((z * torch.tensor([[1.0], [-0.5]], dtype=torch.float64)).sum(), [x] + [sd[k] for k in names])
    assert rel_err(grads[0], g["dx"]) < 1e-5
    for k, gr in zip(names, grads[1:]):
        assert abs(float(gr.norm()) - g["grad_norms"][k]) <= 1e-5 * g["grad_norms"][k], k


def test_losses(golden, oracle):
    import gan_danet_b200 as P
    g = golden("losses_32x64")
    a, b = g["a"].double().requires_grad_(True), g["b"].double()
    tv = oracle.tv_loss(a, 1e-5)
    assert abs(float(tv) - g["tv"]) < 1e-6 * abs(g["tv"])
    (dtv,) = torch.autograd.grad(tv, a)
    assert rel_err(dtv, g["dtv"]) < TOL
    assert abs(float(oracle.ssim(a.detach(), b)) - g["ssim"]) < 1e-6
    assert abs(float(oracle.mse(a.detach(), b)) - g["mse"]) < 1e-6 * g["mse"]
    z = g["z"].double()
    assert abs(float(oracle.bce_with_logits(z, 1.0)) - g["bce1"]) < 1e-7
    assert abs(float(oracle.bce_with_logits(z, 0.0)) - g["bce0"]) < 1e-7
    torch.manual_seed(g["vgg_seed"])
    perc = P.PerceptualLoss(pretrained=False, device=torch.device("cpu"))
    vgg_sd = {k: v.double() for k, v in perc.vgg.state_dict().items()}
    pl = oracle.perceptual_loss(vgg_sd, a, b)
    assert abs(float(pl) - g["perceptual"]) < 1e-6 * abs(g["perceptual"])
    (dpl,) = torch.autograd.grad(pl, a)
    assert rel_err(dpl, g["dperceptual"]) < 1e-5


def test_resample(golden, oracle):
    g = golden("resample")
    x = g["x"].double()
    assert rel_err(oracle.bicubic_up2(x), g["up2"]) < TOL
    assert rel_err(oracle.bilinear_to(x, (20, 28)), g["bil"]) < TOL
    x8 = g["x8"].double()
    assert rel_err(oracle.bicubic_down(x8, 0.5), g["down2"]) < TOL
    assert rel_err(oracle.bicubic_down(x8, 0.25), g["down4"]) < TOL


def test_train_steps(golden, oracle):
    """Two full G+D steps of the oracle vs the reference modules + torch.optim.AdamW (notebook loop)."""
    import gan_danet_b200 as P
    from gan_danet_b200.synthetic import make_batch
    g = golden("train_2steps_8x16")
    lr05, real, aux = make_batch(0, 2, 8, 16)
    torch.manual_seed(g["seed"])
    G = P.FlexibleUpsamplingModule(46)
    D = P.Discriminator1()
    G.apply(P.weights_init_normal)
    for mod in (D.conv1, D.conv2, D.conv3, D.conv4, D.fc2):
        mod.apply(P.weights_init_normal)
    D._materialise_fc1(real)
    torch.manual_seed(g["vgg_seed"])
    perc = P.PerceptualLoss(pretrained=False, device=torch.device("cpu"))
    st = oracle.TrainState(_d(G.state_dict()), _d(D.state_dict()), {k: v.double() for k, v in perc.vgg.state_dict().items()})
    for step in range(2):
        out = oracle.train_step(st, lr05.double(), real.double(), aux.double(), g["epoch"], g["epochs"])
        ref = g["history"][step]
        for k in ("loss_D", "loss_G", "adv", "pixel", "ssim", "tv", "perceptual"):
            assert abs(out[k] - ref[k]) <= 2e-6 * max(abs(ref[k]), 1e-3), (step, k, out[k], ref[k])
    assert rel_err(st.g["final.weight"], g["final_w"]) < TOL
    assert rel_err(st.d["fc2.weight"], g["d_fc2_w"]) < TOL
    assert rel_err(st.g["initial.1.running_mean"], g["initial_bn_rm"]) < TOL


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="reference checkout only exists in the build container")
def test_oracle_vs_live_reference(oracle):
    """Live re-check against the imported reference (build container only)."""
    sys.path.insert(0, "/root/reference")
    try:
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]
        import importlib
        M = importlib.import_module("models")
    finally:
        sys.path.pop(0)
    torch.manual_seed(5)
    G = M.FlexibleUpsamplingModule(40, attention_type="danet")
    G.apply(M.weights_init_normal)
    with torch.no_grad():
        for n, p in G.named_parameters():
            if n.endswith("gamma"):
                p.fill_(0.3)
    G = G.double().train()
    x = torch.randn(2, 40, 6, 10, dtype=torch.float64)
    sd = {k: v.clone() for k, v in G.state_dict().items()}
    y_ref = G(x)
    y = oracle.generator_forward(sd, x, training=True)
    assert rel_err(y, y_ref) < 1e-10


# ------------------------------------------------------------------------------------------------ rows after the step
@pytest.fixture(scope="module")
def post_oracle():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import postprocess_oracle
    return postprocess_oracle


def test_post_hist_match(golden, post_oracle):
    """oracle hist_match == the reference's apply_mild_histogram_matching / simple_histogram_matching (test.ipynb:104-131)
    on continuous, tied and unequal-size samples; the reference forms (1-w)*source in float32, hence 1e-6."""
    import numpy as np
    g = golden("postprocess")
    for case in g["hist"]:
        got = np.stack([post_oracle.hist_match(s, r, case["weight"]) for s, r in zip(case["src"].numpy(), case["ref"].numpy())])
        assert np.abs(got - case["out"].numpy()).max() < 1e-6 * max(1.0, float(case["out"].abs().max())), case["weight"]
    one = g["hist_simple"]
    assert np.abs(post_oracle.hist_match(one["src"].numpy(), one["ref"].numpy(), 1.0) - one["out"].numpy()).max() < 1e-12


def test_post_blend_and_uncertainty(golden, post_oracle):
    import numpy as np
    g = golden("postprocess")
    for case in g["blend"]:
        got = post_oracle.smooth_blend(case["a"].numpy(), case["b"].numpy(), tuple(case["region"]), case["sigma"])
        assert np.abs(got - case["out"].numpy()).max() < 1e-6
    u = g["uncertainty"]
    _, mean_preds, std_preds, r2 = post_oracle.compute_uncertainty(u["preds"].numpy(), u["trues"].numpy(), u["keep"].numpy())
    assert np.abs(mean_preds - u["mean_preds"].numpy()).max() < 1e-6
    assert np.abs(std_preds - u["std_preds"].numpy()).max() < 1e-6
    assert abs(r2 - u["r2"]) < 1e-6          # the reference averages in float32


def test_post_resize(golden, oracle):
    """The oracle's separable resampling matrices at the inference pipeline's scale factors (x1.25, x4: test.ipynb:553,559)."""
    g = golden("postprocess")
    for case in g["resize"]:
        x = case["x"].double()
        ho, wo = case["out"].shape[-2:]
        got = oracle.resample2d(x, (ho, wo), "bicubic", scale=float(case["scale"]))
        assert rel_err(got, case["out"]) < 1e-12, case["scale"]


def test_quantised_oracle_reduces_to_the_reference(golden):
    """oracle/quantised_oracle.py with every rounding point switched off IS the reference generator (float64 golden, 1e-7); with the product
    mode's rounding points on, its distance to the reference is the cost of bf16 operands (~1e-2 on the output), and the gradient
    error follows the sqrt law of a ReLU network (dx ~ sqrt(y error))."""
    from conftest import ROOT
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import quantised_oracle as Q
    import precision_bisect as PB
    g = golden("generator_cin46_8x16")
    exact = PB.summarise(g, *PB.run(g, Q.Formats.exact()))
    assert max(exact.values()) < 1e-7, exact
    prod = PB.summarise(g, *PB.run(g, Q.Formats()))
    assert 2e-3 < prod["y"] < 2e-2, prod
    assert prod["dx"] < 6.0 * prod["y"] ** 0.5, prod
    gonly = PB.summarise(g, *PB.run(g, Q.Formats(None, None, "bf16", None, None, None, None)))
    assert gonly["y"] < 1e-7 and gonly["dx"] < 1e-2, gonly          # rounding only the gradient operands never touches the forward
    # the benchmarked mode (engine.generator_forward_x3): forward convolutions on hi+lo split operands, gradient GEMMs on the bf16 hi parts.  The forward
    # is the parity mode's (<= 1e-3 on the output: north_star) and, with no flipped ReLU masks, the gradients are ~5x closer than with a bf16 forward
    x3 = PB.summarise(g, *PB.run(g, Q.Formats.forward_x3()))
    assert x3["y"] < 3e-4 and x3["dx"] < 6e-2 and x3["grads_whole_vector"] < 6e-2, x3
    assert x3["dx"] < 0.4 * prod["dx"], (x3, prod)
    t = torch.randn(1000, dtype=torch.float64)
    assert float((Q.rnd(t, "bf16x3") - t).abs().max() / t.abs().max()) < 2.0 ** -15          # hi + lo keeps ~16 mantissa bits


def test_notebook_loop_restatement_reproduces_the_golden_trajectory(golden):
    """oracle/notebook_step.py (the notebook's loop GAN_DANet_train.ipynb:182-194,225-269 over a module namespace) driven with the reference's own
    modules in float64 must reproduce the two-step golden trajectory that oracle/make_golden.py recorded from its own inline copy of the loop."""
    from conftest import ROOT
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import notebook_step as NS
    ref_root = next((r for r in (os.path.join(ROOT, "oracle", "_ref"), "/root/reference") if os.path.isdir(os.path.join(r, "models"))), None)
    if ref_root is None:
        pytest.skip("the reference's modules are neither under oracle/_ref nor /root/reference")
    from gan_danet_b200.synthetic import make_batch
    g = golden("train_2steps_8x16")
    M = NS.load_reference_models(ref_root)
    lr05, real, aux = make_batch(0, 2, 8, 16)
    torch.manual_seed(g["seed"])
    G = M.FlexibleUpsamplingModule(46)
    D = M.Discriminator1()
    NS.init_like_the_authors(M, G, D, real)
    torch.manual_seed(g["vgg_seed"])
    perc = M.PerceptualLoss(pretrained=False, device=torch.device("cpu"))
    G, D = G.double().train(), D.double().train()
    perc.vgg.double()
    tr = NS.NotebookTrainer(M, G, D, epochs=g["epochs"], device="cpu", perceptual=perc)
    tr.ssim_loss = tr.ssim_loss.double()
    for step in range(2):
        out = tr.step(lr05.double(), real.double(), aux.double(), epoch=g["epoch"])
        for k, v in g["history"][step].items():
            assert abs(out[k] - v) <= 1e-9 * max(abs(v), 1e-3), (step, k, out[k], v)
    assert rel_err(G.final.weight, g["final_w"]) < 1e-6
