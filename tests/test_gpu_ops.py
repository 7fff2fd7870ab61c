"""torch.ops.gandanet.* (gan_danet_b200/ops.py, SURVEY 8b custom-op layer): the ops with their registered autograd formulas against
the float64 CPU oracle / ATen, through the C ABI.  Tolerances: fp32 engine 1e-5 / 1e-3 (north_star), fused tcgen05 PAM 2e-3 / 1e-2."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

rel = rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module", autouse=True)
def _require_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gan_danet_b200 import _lib
    _lib.lib_for_device(0)
    import gan_danet_b200.ops  # noqa: F401  (registers the ops)


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def _nchw(t):
    return t.permute(0, 3, 1, 2)


@pytest.mark.parametrize("precision,tol_y,tol_g", [("fp32", 1e-5, 1e-3), ("fp16", 2e-3, 2e-2)])
def test_pam_op(golden, oracle, precision, tol_y, tol_g):
    """pam_fwd + its autograd (pam_bwd) on the q/k/v projections of the golden PAM module (generator.py:104-122)."""
    g = golden("pam_c160_8x16")
    sd = {k: v.double() for k, v in g["sd"].items()}
    x = g["x"].double()
    leaves = [oracle.conv2d(x, sd[f"{n}.weight"], sd[f"{n}.bias"]).detach().requires_grad_(True) for n in ("query", "key", "value")]
    xq = x.clone().requires_grad_(True)
    gam = sd["gamma"].clone().requires_grad_(True)
    b, c, h, w = x.shape
    qm, km, vm = (t.reshape(b, -1, h * w) for t in leaves)
    attn = oracle.softmax_lastdim(torch.einsum("bdi,bdj->bij", qm, km))
    ref = gam * torch.einsum("bcj,bij->bci", vm, attn).reshape(b, c, h, w) + xq
    want = torch.autograd.grad((ref * g["r"].double()).sum(), [xq] + leaves + [gam])
    dev_in = [_nhwc(t.detach().float()).to(DEV).requires_grad_(True) for t in [xq] + leaves]
    gd = gam.detach().float().to(DEV).requires_grad_(True)
    y, o, lse = torch.ops.gandanet.pam_fwd(*dev_in, gd, precision)
    y.backward(_nhwc(g["r"]).to(DEV))
    torch.cuda.synchronize()
    assert rel_err(_nchw(y), ref) < tol_y, rel_err(_nchw(y), ref)
    assert o.shape == (b, h * w, c) and lse.shape == (b, h * w)
    for got, w_ in zip([t.grad for t in dev_in], want[:4]):
        assert rel_err(_nchw(got), w_) < tol_g, rel_err(_nchw(got), w_)
    assert rel_err(gd.grad, want[4]) < tol_g


@pytest.mark.parametrize("tc,tol_y,tol_g", [(False, 1e-5, 1e-3), (True, 1e-4, 5e-3)])
def test_cam_op(golden, tc, tol_y, tol_g):
    g = golden("cam_c160_8x16")
    x = _nhwc(g["x"]).to(DEV).requires_grad_(True)
    gam = g["sd"]["gamma"].to(DEV).requires_grad_(True)
    y, attn = torch.ops.gandanet.cam_fwd(x, gam, tc)
    y.backward(_nhwc(g["r"]).to(DEV))
    assert rel_err(_nchw(y), g["y"]) < tol_y
    assert rel_err(_nchw(x.grad), g["dx"]) < tol_g
    assert rel_err(gam.grad, g["grads"]["gamma"]) < tol_g
    assert attn.shape == (x.shape[0], 160, 160) and float((attn.sum(-1) - 1).abs().max()) < 1e-4


@pytest.mark.parametrize("prec,tol", [("fp32", 1e-5), ("bf16x3", 1e-4), ("bf16", 2e-2)])
@pytest.mark.parametrize("cin,cout,k,stride,pad,act", [(46, 64, 3, 1, 1, 1), (160, 80, 1, 1, 0, 0), (64, 128, 3, 2, 1, 2)])
def test_conv2d_op(prec, tol, cin, cout, k, stride, pad, act):
    """conv2d + fused activation + autograd (conv2d_bwd) against ATen in float64 on every engine."""
    from gan_danet_b200 import engine as E
    gen = torch.Generator().manual_seed(cin)
    x = torch.randn(2, cin, 16, 24, generator=gen)
    w = torch.randn(cout, cin, k, k, generator=gen) * (cin * k * k) ** -0.5
    bias = torch.randn(cout, generator=gen) * 0.1
    xd, wd, bd = (t.double().requires_grad_(True) for t in (x, w, bias))
    z = F.conv2d(xd, wd, bd, stride=stride, padding=pad)
    ref = z if act == 0 else (F.relu(z) if act == 1 else F.leaky_relu(z, 0.2))
    r = torch.randn(ref.shape, generator=gen)
    want = torch.autograd.grad((ref * r.double()).sum(), [xd, wd, bd])
    old = E.conv_precision
    E.set_conv_precision(prec)
    try:
        xg = _nhwc(x).to(DEV).requires_grad_(True)
        wg, bg = w.to(DEV).requires_grad_(True), bias.to(DEV).requires_grad_(True)
        y = torch.ops.gandanet.conv2d(xg, wg, bg, stride, pad, act, 0.2)
        y.backward(_nhwc(r).to(DEV))
        torch.cuda.synchronize()
    finally:
        E.set_conv_precision(old)
    assert rel_err(_nchw(y), ref) < tol, rel_err(_nchw(y), ref)
    assert rel_err(_nchw(xg.grad), want[0]) < 3 * tol and rel_err(wg.grad, want[1]) < 3 * tol and rel_err(bg.grad, want[2]) < 3 * tol


def test_bicubic_and_adamw_ops():
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(2, 8, 9, 13, generator=gen)
    xd = x.double().requires_grad_(True)
    ref = F.interpolate(xd, scale_factor=2, mode="bicubic", align_corners=False)
    r = torch.randn(ref.shape, generator=gen)
    (want,) = torch.autograd.grad((ref * r.double()).sum(), [xd])
    xg = _nhwc(x).to(DEV).requires_grad_(True)
    y = torch.ops.gandanet.upsample_bicubic2x(xg)
    y.backward(_nhwc(r).to(DEV))
    assert rel_err(_nchw(y), ref) < 1e-6 and rel_err(_nchw(xg.grad), want) < 1e-6
    # fused_adamw_ against torch.optim.AdamW (GAN_DANet_train.ipynb:182-183 hyper-parameters), three steps
    p0 = torch.randn(1000, generator=gen)
    grads = [torch.randn(1000, generator=gen) for _ in range(3)]
    pr = torch.nn.Parameter(p0.clone().double())
    opt = torch.optim.AdamW([pr], lr=2e-4, betas=(0.5, 0.999), weight_decay=1e-4)
    p, m, v = p0.clone().to(DEV), torch.zeros(1000, device=DEV), torch.zeros(1000, device=DEV)
    for step, gr in enumerate(grads, 1):
        pr.grad = gr.double()
        opt.step()
        torch.ops.gandanet.fused_adamw_(p, gr.to(DEV), m, v, 2e-4, 0.5, 0.999, 1e-8, 1e-4, step, 1.0)
    assert rel_err(p, pr) < 1e-6


def test_python_free_cabi_harness():
    """tools/cabi_bench.cpp (built by gan_danet_b200/build.py): plain C++ over the C ABI, no torch -- runs the fused PAM forward and
    backward and checks sampled rows against its own float64 host evaluation of generator.py:115-122 (exit code 0 = within 2e-3 / 6e-3)."""
    import json
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gan_danet_b200", "cabi_bench")
    if not os.path.exists(exe):
        from gan_danet_b200.build import build_harness
        build_harness()
    r = subprocess.run([exe, "2", "2048", "184", "23"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout, r.stderr)
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["rel_err_y"] < 2e-3 and out["pam_fwd_tflops"] > 0


def test_bn_stats_finalize_op():
    """torch.ops.gandanet.bn_stats_finalize against nn.BatchNorm2d in train mode: normalisation coefficients and the in-place running statistics."""
    import gan_danet_b200.ops  # noqa: F401
    g = torch.Generator().manual_seed(3)
    x = (2.0 * torch.randn(3, 5, 7, 24, generator=g) + 0.7).to(DEV)                  # NHWC
    bn = torch.nn.BatchNorm2d(24).double()
    with torch.no_grad():
        bn.weight.copy_(torch.rand(24, generator=g) + 0.5); bn.bias.copy_(torch.randn(24, generator=g))
    w, b = bn.weight.detach().float().to(DEV), bn.bias.detach().float().to(DEV)
    rm, rv = torch.zeros(24, device=DEV), torch.ones(24, device=DEV)
    coef = torch.ops.gandanet.bn_stats_finalize(x, w, b, rm, rv, 1e-5, 0.1)
    y = x * coef[2] + coef[3]
    ref = bn.train()(x.cpu().double().permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
    assert rel(y, ref) < 1e-5
    assert rel(rm, bn.running_mean) < 1e-5 and rel(rv, bn.running_var) < 1e-5


def test_bilinear_resize_add_op():
    import torch.nn.functional as F
    import gan_danet_b200.ops  # noqa: F401
    g = torch.Generator().manual_seed(4)
    s = torch.randn(2, 6, 9, 8, generator=g).to(DEV).requires_grad_(True)
    x = torch.randn(2, 24, 36, 8, generator=g).to(DEV).requires_grad_(True)
    r = torch.randn(2, 24, 36, 8, generator=g).to(DEV)
    y = torch.ops.gandanet.bilinear_resize_add_fwd(s, x)
    (y * r).sum().backward()
    sd, xd = s.detach().cpu().double().requires_grad_(True), x.detach().cpu().double().requires_grad_(True)
    ref = xd + F.interpolate(sd.permute(0, 3, 1, 2), size=(24, 36), mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
    (ref * r.cpu().double()).sum().backward()
    assert rel(y, ref) < 1e-6 and rel(s.grad, sd.grad) < 1e-5 and rel(x.grad, xd.grad) < 1e-6


def test_multi_loss_op():
    """torch.ops.gandanet.multi_loss_fwd_bwd against nn.MSELoss / nn.BCEWithLogitsLoss / the reference's TVLoss formula (losses.py:81-87) and autograd."""
    import gan_danet_b200.ops  # noqa: F401
    g = torch.Generator().manual_seed(5)
    hr = torch.randn(3, 1, 16, 24, generator=g).to(DEV)
    real = torch.randn(3, 1, 16, 24, generator=g).to(DEV)
    z = torch.randn(3, 1, generator=g).to(DEV)
    w, tvw = 0.02, 1e-5
    losses, dhr, dz = torch.ops.gandanet.multi_loss_fwd_bwd(hr, real, z, w, tvw)
    h, zz = hr.cpu().double().requires_grad_(True), z.cpu().double().requires_grad_(True)
    pix = torch.nn.functional.mse_loss(h, real.cpu().double())
    adv = torch.nn.functional.binary_cross_entropy_with_logits(zz, torch.ones_like(zz))
    B, Cc, H, W = h.shape
    tv = tvw * 2 * (((h[:, :, 1:] - h[:, :, :-1]) ** 2).sum() / (B * Cc * (H - 1) * W) + ((h[:, :, :, 1:] - h[:, :, :, :-1]) ** 2).sum() / (B * Cc * H * (W - 1))) / B
    total = (1 - w) * pix + w * adv + tv
    total.backward()
    got = losses.cpu().double()
    for a, b in zip(got, (total, pix, adv, tv)):
        assert abs(float(a) - float(b)) < 1e-5 * abs(float(b)) + 1e-12, (float(a), float(b))
    assert rel(dhr, h.grad) < 1e-5 and rel(dz, zz.grad) < 1e-5
