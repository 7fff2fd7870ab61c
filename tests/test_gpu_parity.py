"""Parity of the CUDA path (through the module API -> C ABI) against the CPU oracle and the committed golden vectors.

Error metric: ||a-b||_2/||b||_2 against float64 references.  Tolerances (stated per test):
  fp32 CUDA-core engine: 1e-4 on outputs, 1e-3 on gradients (north_star: "within 1e-3 relative error under fp32 accumulation");
  fp16-operand tcgen05 PAM kernel: 2e-3 on y at gamma = 0.5 (SURVEY appendix C measures 8.9e-4 for fp16 operands).
"""
import os

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module", autouse=True)
def _require_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gan_danet_b200 import _lib
    _lib.lib_for_device(0)     # raises loudly if the library is missing or the device is not sm_100


@pytest.fixture(autouse=True)
def _fp32_engine_by_default():
    """The parity tests of this file run on the fp32 engine unless a test selects a tensor-core precision itself."""
    from gan_danet_b200 import engine as E
    old, old_cam = E.conv_precision, E.cam_tensor_core
    E.set_conv_precision("fp32")
    E.cam_tensor_core = None
    yield
    E.set_conv_precision(old)
    E.cam_tensor_core = old_cam


def _to(sd, dev=DEV):
    return {k: v.to(dev) for k, v in sd.items()}


def _fwd_bwd(mod, x, r):
    mod = mod.to(DEV)
    for p in mod.parameters():
        p.grad = None
    xg = x.to(DEV).requires_grad_(True)
    y = mod(xg)
    y.backward(r.to(DEV))
    torch.cuda.synchronize()
    return y.detach(), xg.grad.detach(), {k: p.grad.detach() for k, p in mod.named_parameters() if p.grad is not None}


# ------------------------------------------------------------------------------------------------ primitive kernels


@pytest.mark.parametrize("cin,cout,k,stride,pad,hw", [(46, 64, 3, 1, 1, (8, 16)), (64, 24, 3, 1, 1, (9, 7)), (160, 80, 1, 1, 0, (8, 16)),
                                                    (1, 64, 3, 2, 1, (32, 64)), (64, 128, 3, 2, 1, (15, 23)), (320, 160, 3, 1, 1, (8, 8))])
def test_conv_forward_backward(cin, cout, k, stride, pad, hw):
    """nn.Conv2d semantics (generator.py:20-228, discriminator.py:62-65): forward, data gradient, weight/bias gradient."""
    from gan_danet_b200 import engine as E
    from gan_danet_b200.models.generator import TapeModule, _conv

    class M(TapeModule):
        def __init__(self):
            super().__init__()
            self.c = torch.nn.Conv2d(cin, cout, k, stride=stride, padding=pad)

        def _build(self, ctx, x):
            return _conv(ctx, x, self.c, act=2, slope=0.2)

    torch.manual_seed(0)
    m = M()
    x = torch.randn(3, cin, *hw)
    xd = x.double().requires_grad_(True)
    yref = torch.nn.functional.leaky_relu(torch.nn.functional.conv2d(xd, m.c.weight.double(), m.c.bias.double(), stride=stride, padding=pad), 0.2)
    r = torch.randn(yref.shape)
    wd, bd = m.c.weight.detach().double().requires_grad_(True), m.c.bias.detach().double().requires_grad_(True)
    y2 = torch.nn.functional.leaky_relu(torch.nn.functional.conv2d(xd, wd, bd, stride=stride, padding=pad), 0.2)
    gx, gw, gb = torch.autograd.grad((y2 * r.double()).sum(), [xd, wd, bd])
    y, dx, grads = _fwd_bwd(m, x, r)
    assert rel_err(y, yref) < 1e-5
    assert rel_err(dx, gx) < 1e-5
    assert rel_err(grads["c.weight"], gw) < 1e-5
    assert rel_err(grads["c.bias"], gb) < 1e-5


def test_resample_kernels(golden, oracle):
    """bicubic x2 (generator.py:221,225), bilinear (generator.py:244) and the bicubic input down-sampling
    (GAN_DANet_train.ipynb:226,231) against ATen outputs stored in the golden fixture; backward against the oracle."""
    from gan_danet_b200 import engine as E
    from gan_danet_b200.trainer import prepare_input_nhwc
    g = golden("resample")
    x = g["x"]                                            # [1,3,5,7]
    xn = x.permute(0, 2, 3, 1).contiguous().to(DEV)
    tape = E.Tape()
    xv = E.Var(xn)
    up = E.op_bicubic_up2(tape, xv)
    assert rel_err(up.t.permute(0, 3, 1, 2), g["up2"]) < 1e-6
    r = torch.randn(up.t.shape, generator=torch.Generator().manual_seed(1))
    up.g = r.to(DEV)
    tape.backward()
    xd = x.double().requires_grad_(True)
    (gx,) = torch.autograd.grad((oracle.bicubic_up2(xd) * r.permute(0, 3, 1, 2).double()).sum(), xd)
    assert rel_err(xv.g.permute(0, 3, 1, 2), gx) < 1e-6
    # bilinear add
    tape = E.Tape()
    sv = E.Var(xn.clone())
    base = torch.zeros(1, 20, 28, 3, device=DEV)
    out = E.op_bilinear_add_(tape, sv, E.Var(base))
    assert rel_err(out.t.permute(0, 3, 1, 2), g["bil"]) < 1e-6
    r2 = torch.randn(out.t.shape, generator=torch.Generator().manual_seed(2))
    out.g = r2.to(DEV)
    tape.backward()
    (gb,) = torch.autograd.grad((oracle.bilinear_to(xd, (20, 28)) * r2.permute(0, 3, 1, 2).double()).sum(), xd)
    assert rel_err(sv.g.permute(0, 3, 1, 2), gb) < 1e-6
    # input preparation: [B,1,2h,2w] and [B,Ca,4h,4w] -> NHWC [B,h,w,1+Ca]
    lr05 = torch.randn(2, 1, 8, 12, generator=torch.Generator().manual_seed(3))
    aux = torch.randn(2, 5, 16, 24, generator=torch.Generator().manual_seed(4))
    xin = prepare_input_nhwc(lr05.to(DEV), aux.to(DEV))
    ref = oracle.prepare_input(lr05.double(), aux.double())
    assert rel_err(xin.permute(0, 3, 1, 2), ref) < 1e-6
    assert rel_err(prepare_input_nhwc(g["x8"].to(DEV), torch.zeros(1, 1, 16, 24, device=DEV))[..., :1].permute(0, 3, 1, 2), g["down2"][:, :1]) < 1e-6


def test_losses(golden, oracle):
    """TV / SSIM / MSE / BCE / perceptual (losses.py:13-147, GAN_DANet_train.ipynb:190-194) values and gradients."""
    import gan_danet_b200 as P
    g = golden("losses_32x64")
    a = g["a"].to(DEV).requires_grad_(True)
    b = g["b"].to(DEV)
    tv = P.TVLoss(1e-5)(a)
    (dtv,) = torch.autograd.grad(tv, a)
    assert abs(float(tv) - g["tv"]) < 1e-5 * abs(g["tv"])
    assert rel_err(dtv, g["dtv"]) < 1e-5
    assert abs(float(P.SSIM().to(DEV)(a.detach(), b)) - g["ssim"]) < 1e-5
    mse = P.MSELoss()(a, b)
    (dm,) = torch.autograd.grad(3.0 * mse, a)
    assert abs(float(mse) - g["mse"]) < 1e-5 * g["mse"]
    assert rel_err(dm, 3.0 * 2.0 * (g["a"] - g["b"]).double() / g["a"].numel()) < 1e-5
    z = g["z"].to(DEV).requires_grad_(True)
    l1 = P.BCEWithLogitsLoss()(z, torch.ones_like(z))
    l0 = P.BCEWithLogitsLoss()(z, torch.zeros_like(z))
    assert abs(float(l1) - g["bce1"]) < 1e-6 and abs(float(l0) - g["bce0"]) < 1e-6
    (dz,) = torch.autograd.grad(l1, z)
    assert rel_err(dz, (torch.sigmoid(g["z"].double()) - 1.0) / 6.0) < 1e-5
    torch.manual_seed(g["vgg_seed"])
    perc = P.PerceptualLoss(pretrained=False, device=torch.device("cpu"))
    perc.vgg.to(DEV)
    perc.device = torch.device(DEV)
    pl = perc(a, b)
    (dpl,) = torch.autograd.grad(pl, a)
    assert abs(float(pl) - g["perceptual"]) < 1e-4 * abs(g["perceptual"])
    assert rel_err(dpl, g["dperceptual"]) < 1e-3


def test_fused_adamw_matches_torch():
    """torch.optim.AdamW (GAN_DANet_train.ipynb:182-183) vs gdn_adamw over 5 steps."""
    from gan_danet_b200.trainer import FusedAdamW
    torch.manual_seed(0)
    p0 = torch.randn(1000, 37)
    gs = [torch.randn(1000, 37) for _ in range(5)]
    pr = torch.nn.Parameter(p0.clone().double())
    opt_r = torch.optim.AdamW([pr], lr=4e-4, betas=(0.5, 0.999), weight_decay=1e-4)
    pg = torch.nn.Parameter(p0.clone().to(DEV))
    opt_g = FusedAdamW([pg], lr=4e-4, betas=(0.5, 0.999), weight_decay=1e-4)
    for g in gs:
        pr.grad = g.double()
        opt_r.step()
        pg.grad = g.to(DEV)
        opt_g.step()
    assert rel_err(pg, pr) < 1e-6
    assert rel_err(pg.detach().cpu().double() - p0.double(), pr.detach() - p0.double()) < 1e-3


# ------------------------------------------------------------------------------------------------ modules vs golden


def _check_grads(grads, gold, tol, skip=()):
    for k, ref in gold.items():
        if any(s in k for s in skip):
            continue
        assert k in grads, k
        assert rel_err(grads[k], ref) < tol, (k, rel_err(grads[k], ref))


@pytest.mark.parametrize("name,precision,tol_y,tol_g", [("pam_c160_8x16", "fp32", 1e-5, 1e-4), ("pam_c184_4x8", "fp32", 1e-5, 1e-4),
                                                       ("pam_c160_8x16", "fp16", 2e-3, 1e-2), ("pam_c160_8x16", "fp16x3", 1e-3, 5e-3),
                                                       ("pam_c184_4x8", "fp16x3", 1e-3, 5e-3)])
def test_pam_module(golden, name, precision, tol_y, tol_g):
    """PAMModule (generator.py:104-122), gamma = 0.5, against the reference's float64 run.  fp16 / fp16x3 = fused tcgen05 flash forward and
    backward with single / hi+lo split fp16 logit operands (the 4x8 grid, N = 32, goes through the padded path)."""
    from gan_danet_b200.models.generator import PAMModule
    g = golden(name)
    C = g["x"].shape[1]
    m = PAMModule(C)
    m.load_state_dict(g["sd"])
    m.precision = precision
    y, dx, grads = _fwd_bwd(m, g["x"], g["r"])
    assert rel_err(y, g["y"]) < tol_y, rel_err(y, g["y"])
    assert rel_err(dx, g["dx"]) < tol_g, rel_err(dx, g["dx"])
    _check_grads(grads, g["grads"], tol_g, skip=("key.bias",))
    # d/d(key.bias) is analytically zero: compare against the weight-gradient scale (the tensor-core backward rounds dS to
    # bf16, so its column sums cancel to 2^-9 of the terms instead of fp32 epsilon)
    assert grads["key.bias"].abs().max() < (1e-3 if precision == "fp32" else 1e-2) * grads["key.weight"].abs().max() + 1e-6, \
        (float(grads["key.bias"].abs().max()), float(grads["key.weight"].abs().max()))


@pytest.mark.parametrize("B,C,hw", [(2, 184, (32, 32)), (1, 160, (64, 128))])
def test_pam_flash_kernel_vs_oracle(oracle, B, C, hw):
    """The fused tcgen05 kernel at N = 1024 and at the north-star N = 8192 against the float64 oracle (blocked PAM)."""
    from gan_danet_b200.models.generator import PAMModule
    import gan_danet_b200 as P
    torch.manual_seed(1)
    m = PAMModule(C)
    m.apply(P.weights_init_normal)
    with torch.no_grad():
        m.gamma.fill_(0.5)
    x = 0.5 * torch.randn(B, C, *hw)
    sd = {k: v.double() for k, v in m.state_dict().items()}
    ref = oracle.pam_blocked(x.double(), sd["query.weight"], sd["query.bias"], sd["key.weight"], sd["key.bias"], sd["value.weight"], sd["value.bias"],
                             sd["gamma"], block=512)
    m.precision = "fp16"
    m = m.to(DEV)
    with torch.no_grad():
        y16 = m(x.to(DEV))
        m.precision = "fp16x3"
        y16x3 = m(x.to(DEV))
        m.precision = "fp32"
        y32 = m(x.to(DEV))
    torch.cuda.synchronize()
    assert rel_err(y32, ref) < 1e-5, rel_err(y32, ref)
    out_ref = (ref - x.double()) / 0.5
    e16 = rel_err((y16.cpu().double() - x.double()) / 0.5, out_ref)
    assert e16 < 4e-3, e16                        # attention term itself (1.0-1.2e-3 at N = 8192 for fp16 operands)
    assert rel_err(y16, ref) < 2e-3
    e16x3 = rel_err((y16x3.cpu().double() - x.double()) / 0.5, out_ref)
    assert e16x3 < 3e-3 and e16x3 <= e16 * 1.05, (e16x3, e16)      # bf16 P and V remain (1.8e-3 at reference-scale logits)
    assert rel_err(y16x3, ref) < 1e-3, rel_err(y16x3, ref)


@pytest.mark.parametrize("B,C,hw", [(2, 184, (45, 22)), (1, 160, (9, 13)), (2, 176, (4, 8))])
def test_pam_padded_grid_vs_oracle(oracle, B, C, hw):
    """Grids whose N is not a multiple of the kernels' 128-row tiles (the authors' 45x22 grid: N = 990) run on the tensor-core
    kernels padded per sample, padded keys masked through a spare logit column (engine._op_pam_core_padded): forward, dx and all
    parameter gradients against the float64 oracle at the fp16-kernel tolerances, and the padded softmax must not leak
    (constant value map => gamma*v + x exactly)."""
    from gan_danet_b200.models.generator import PAMModule
    from gan_danet_b200 import engine as E
    from gan_danet_b200._lib import PREC_FP16
    import gan_danet_b200 as P
    torch.manual_seed(2)
    m = PAMModule(C)
    m.apply(P.weights_init_normal)
    with torch.no_grad():
        m.gamma.fill_(0.5)
        m.query.weight.mul_(4.0)
        m.key.weight.mul_(4.0)
    gen = torch.Generator().manual_seed(3)
    x = 0.5 * torch.randn(B, C, *hw, generator=gen)
    r = torch.randn(B, C, *hw, generator=gen)
    sd = {k: v.detach().double().requires_grad_(True) for k, v in m.state_dict().items()}
    xd = x.double().requires_grad_(True)
    ref = oracle.pam(xd, sd["query.weight"], sd["query.bias"], sd["key.weight"], sd["key.bias"], sd["value.weight"], sd["value.bias"], sd["gamma"])
    names = ["query.weight", "query.bias", "key.weight", "value.weight", "value.bias", "gamma"]
    want = torch.autograd.grad((ref * r.double()).sum(), [xd] + [sd[n] for n in names])
    m.precision = "fp16x3"
    assert E.pam_pad_to_tiles
    y, dx, grads = _fwd_bwd(m, x, r)
    assert rel_err(y, ref) < 1e-3, rel_err(y, ref)
    assert rel_err((y.cpu().double() - x.double()), (ref.detach() - x.double())) < 4e-3
    # gradients with 4x query/key weights (logits far beyond the reference's +-100): round 1 (single fp16 logits, bf16 gradient operands) measured
    # 1.2e-2 at N = 990 and 3.0e-2 at N = 32 and asserted 5e-2; split logits + fp16 gradient operands: asserted 5e-3 / 1e-2 (N = 32: no averaging over keys)
    tol = 5e-3 if hw[0] * hw[1] >= 128 else 1e-2
    errs = {"dx": rel_err(dx, want[0]), **{n: rel_err(grads[n], w) for n, w in zip(names, want[1:])}}
    assert max(errs.values()) < tol, errs
    H, W = hw
    d = C // 8
    xx = torch.randn(B, H, W, C, generator=gen).to(DEV)
    q = (2.0 * torch.randn(B, H, W, d, generator=gen)).to(DEV)
    k = (2.0 * torch.randn(B, H, W, d, generator=gen)).to(DEV)
    gamma = torch.full((1,), 0.5, device=DEV)
    t = E.Tape(record=False)
    yc = E.op_pam_core(t, E.Var(xx), E.Var(q), E.Var(k), E.Var(torch.full((B, H, W, C), 0.75, device=DEV)), E.Var(gamma), precision=E.PREC_FP16X3).t
    assert float((yc - (xx + 0.5 * 0.75)).abs().max()) < 2e-3


def test_pam_value_operand_from_conv_epilogue():
    """PAMModule in the product mode (bf16 tensor-core convs + fused fp16/bf16 PAM): the value projection's epilogue emits the bf16 V
    operand of the fused kernel (engine.pam_v16_buffer, gdn_pam_fwd_args.v16) instead of a packing pass over V -- the same
    round-to-nearest of the same fp32 values, so forward and every gradient are bitwise those of the packing path."""
    from gan_danet_b200 import engine as E
    from gan_danet_b200.models.generator import PAMModule
    import gan_danet_b200 as P
    gen = torch.Generator().manual_seed(9)
    x = 0.5 * torch.randn(2, 184, 16, 32, generator=gen)
    r = torch.randn(2, 184, 16, 32, generator=gen)

    def run(flag):
        torch.manual_seed(4)
        m = PAMModule(184)
        m.apply(P.weights_init_normal)
        with torch.no_grad():
            m.gamma.fill_(0.5)
        m.precision = "fp16"
        old, oldf = E.conv_precision, E.pam_v16_from_conv
        E.set_conv_precision("bf16")
        E.pam_v16_from_conv = flag
        try:
            assert (E.pam_v16_buffer(torch.empty(2, 16, 32, 184, device=DEV)) is not None) == flag
            return _fwd_bwd(m, x, r)
        finally:
            E.set_conv_precision(old)
            E.pam_v16_from_conv = oldf

    y1, dx1, g1 = run(True)
    y0, dx0, g0 = run(False)
    assert torch.equal(y1, y0) and torch.equal(dx1, dx0)
    assert all(torch.equal(g1[k], g0[k]) for k in g0)
    buf = E.pam_v16_buffer(torch.empty(2, 16, 32, 184, device=DEV)) if E.conv_precision == "bf16" else None
    E.set_conv_precision("bf16")
    try:
        buf = E.pam_v16_buffer(torch.empty(2, 16, 32, 184, device=DEV))
        assert float(buf[:, 184].float().min()) == 1.0 and float(buf[:, 185:].float().abs().max()) == 0.0      # the tail survived the epilogue writes
    finally:
        E.set_conv_precision("fp32")


@pytest.mark.parametrize("conv", ["bf16", "bf16x3"])
@pytest.mark.parametrize("C,hw", [(184, (16, 32)), (160, (8, 16))])
def test_pam_merged_query_key_projection(conv, C, hw):
    """Query and key projections (generator.py:108-109) as one C -> 2d tensor-core convolution with concatenated weights against the two
    separate convolutions: per output channel the same dot products, so y is bitwise equal; gradients differ only in float32 summation order
    (the q and k contributions to dL/dx are added inside one GEMM instead of two accumulation passes)."""
    from gan_danet_b200 import engine as E
    from gan_danet_b200.models.generator import PAMModule
    import gan_danet_b200 as P
    gen = torch.Generator().manual_seed(C)
    x = 0.5 * torch.randn(2, C, *hw, generator=gen)
    r = torch.randn(2, C, *hw, generator=gen)

    def run(flag):
        torch.manual_seed(4)
        m = PAMModule(C)
        m.apply(P.weights_init_normal)
        with torch.no_grad():
            m.gamma.fill_(0.5)
            m.query.weight.mul_(4.0)
            m.key.weight.mul_(4.0)
        m.precision = "fp16"
        old, oldf = E.conv_precision, E.pam_merge_qk
        E.set_conv_precision(conv)
        E.pam_merge_qk = flag
        try:
            return _fwd_bwd(m, x, r)
        finally:
            E.set_conv_precision(old)
            E.pam_merge_qk = oldf

    y1, dx1, g1 = run(True)
    y0, dx0, g0 = run(False)
    assert torch.equal(y1, y0)
    assert rel_err(dx1, dx0) < 1e-5, rel_err(dx1, dx0)
    for k in g0:
        if k == "key.bias":
            assert float((g1[k] - g0[k]).abs().max()) < 1e-4 * float(g0["key.weight"].abs().max()) + 1e-7
        else:
            assert rel_err(g1[k], g0[k]) < 1e-4, (k, rel_err(g1[k], g0[k]))


def test_danet_outputs_written_as_fuse_operand():
    """DANetAttention (generator.py:142-157) in the product mode: PAM and CAM write their results only as bf16 column blocks of the fuse
    convolution's packed operand (gdn_pam_fwd_args.y16, gdn_cam_fwd_tc16) instead of an fp32 cat tensor that is packed afterwards -- the same
    rounding of the same values: output, input gradient and all parameter gradients bitwise equal to the unfused launches."""
    from gan_danet_b200 import engine as E
    from gan_danet_b200.models.generator import DANetAttention
    import gan_danet_b200 as P
    gen = torch.Generator().manual_seed(13)
    x = 0.5 * torch.randn(2, 160, 16, 32, generator=gen)
    r = torch.randn(2, 160, 16, 32, generator=gen)

    def run(flag):
        torch.manual_seed(6)
        m = DANetAttention(160)
        m.apply(P.weights_init_normal)
        with torch.no_grad():
            m.position_attention.gamma.fill_(0.5)
            m.channel_attention.gamma.fill_(0.05)
        m.position_attention.precision = "fp16"
        old, oldf = E.conv_precision, E.danet_cat16
        E.set_conv_precision("bf16")
        E.danet_cat16 = flag
        try:
            assert E.danet_cat16_ok(torch.empty(2, 16, 32, 160, device=DEV), True, 160) == flag      # the test sets the flag itself (default: off)
            return _fwd_bwd(m, x, r)
        finally:
            E.set_conv_precision(old)
            E.danet_cat16 = oldf

    y1, dx1, g1 = run(True)
    y0, dx0, g0 = run(False)
    assert torch.equal(y1, y0) and torch.equal(dx1, dx0)
    bad = [k for k in g0 if not torch.equal(g1[k], g0[k])]
    assert not bad, bad


@pytest.mark.parametrize("prec_name", ["fp16x3", "fp16"])
def test_pam_padding_is_exact(prec_name):
    """The same aligned problem (N = 256) through the kernels directly and through the padded path forced to 512 rows: the
    padded keys get softmax weight 0 (2^-125 on the polynomial lanes) and the padded queries a zero cotangent, so forward and
    all gradients agree to float32 summation-order level -- far below the fp16/bf16 operand rounding."""
    from gan_danet_b200 import engine as E
    prec = E.PAM_PRECISION_NAMES[prec_name]
    B, H, W, C, d = 2, 16, 16, 184, 23
    gen = torch.Generator().manual_seed(11)
    x, v, dy = (torch.randn(B, H, W, C, generator=gen).to(DEV) for _ in range(3))
    q, k = ((1.5 * torch.randn(B, H, W, d, generator=gen)).to(DEV) for _ in range(2))
    gamma = torch.full((1,), 0.5, device=DEV)

    def run(padded):
        tape = E.Tape()
        vs = [E.Var(t) for t in (x, q, k, v)]
        gv = E.Var(gamma)
        y = E._op_pam_core_padded(tape, *vs, gv, None, pad_to=512, precision=prec) if padded else E.op_pam_core(tape, *vs, gv, precision=prec)
        y.g = dy.clone()
        tape.backward()
        torch.cuda.synchronize()
        return [y.t] + [t.g for t in vs] + [gv.g]

    for a, b in zip(run(False), run(True)):
        assert rel_err(b, a) < 2e-6, rel_err(b, a)


@pytest.mark.parametrize("tc,tol_y,tol_g", [(False, 1e-5, 1e-3), (True, 1e-4, 5e-3)])
@pytest.mark.parametrize("name", ["cam_c160_8x16", "cam_c184_4x8"])
def test_cam_module(golden, name, tc, tol_y, tol_g):
    """CAMModule (generator.py:125-139), gamma = 0.5: fp32 CUDA-core engine and the tensor-core path (bf16 hi+lo split
    operands: Gram matrices by the grouped tcgen05 weight-gradient kernel, re-projections by the grouped 1x1 conv)."""
    from gan_danet_b200 import engine as E
    from gan_danet_b200.models.generator import CAMModule
    g = golden(name)
    m = CAMModule(g["x"].shape[1])
    m.load_state_dict(g["sd"])
    old = E.cam_tensor_core
    E.cam_tensor_core = tc
    try:
        y, dx, grads = _fwd_bwd(m, g["x"], g["r"])
    finally:
        E.cam_tensor_core = old
    assert rel_err(y, g["y"]) < tol_y, rel_err(y, g["y"])
    assert rel_err(dx, g["dx"]) < tol_g, rel_err(dx, g["dx"])
    assert rel_err(grads["gamma"], g["grads"]["gamma"]) < tol_g, rel_err(grads["gamma"], g["grads"]["gamma"])


def test_dense_transition_danet(golden):
    import gan_danet_b200 as P
    from gan_danet_b200.models import generator as PG
    g = golden("denseblock_64_8x16")
    m = PG.DenseBlock(4, 64, 24)
    m.load_state_dict(g["sd"])
    m.train()
    y, dx, grads = _fwd_bwd(m, g["x"], g["r"])
    assert rel_err(y, g["y"]) < 1e-5
    assert rel_err(dx, g["dx"]) < 1e-3
    _check_grads(grads, g["grads"], 2e-3)
    g = golden("transition_160_8x16")
    m = PG.TransitionLayer(160, 80)
    m.load_state_dict(g["sd"])
    m.train()
    y, dx, grads = _fwd_bwd(m, g["x"], g["r"])
    assert rel_err(y, g["y"]) < 1e-5 and rel_err(dx, g["dx"]) < 1e-3
    _check_grads(grads, g["grads"], 2e-3)
    g = golden("danet_c160_8x16")
    torch.manual_seed(0)
    m = PG.DANetAttention(160)
    m.apply(P.weights_init_normal)
    m.load_state_dict(g["sd"], strict=False)
    m.position_attention.precision = "fp32"
    m.train()
    y, dx, grads = _fwd_bwd(m, g["x"], g["r"])
    assert rel_err(y, g["y"]) < 1e-4, rel_err(y, g["y"])
    assert rel_err(dx, g["dx"]) < 1e-3, rel_err(dx, g["dx"])
    _check_grads(grads, g["grads"], 2e-3, skip=("key.bias",))


def _make_generator(seed, gamma, precision):
    import gan_danet_b200 as P
    from gan_danet_b200.models.generator import CAMModule, PAMModule
    torch.manual_seed(seed)
    G = P.FlexibleUpsamplingModule(46)
    G.apply(P.weights_init_normal)
    with torch.no_grad():
        for m in G.modules():
            if isinstance(m, (PAMModule, CAMModule)):
                m.gamma.fill_(gamma)
    G.set_pam_precision(precision)
    return G.train()


@pytest.mark.parametrize("conv,precision,tol_y,tol_g", [("fp32", "fp32", 1e-4, 2e-3), ("bf16x3", "fp32", 1e-3, 1e-2), ("fp32", "fp16", 1e-3, 5e-2),
                                                        ("fp32", "fp16x3", 1e-3, 4e-2), ("bf16x3", "fp16x3", 1e-3, 4e-2)])
def test_generator(golden, conv, precision, tol_y, tol_g):
    """FlexibleUpsamplingModule (generator.py:175-247) at C_in 46, grid 8x16, gamma 0.05: output, input gradient,
    per-tensor gradient norms, small gradient tensors and BN running statistics against the reference's float64 run.
    Measured (tools/measure_precision.py, B200): fp32 engine y 8e-7 / dx 4e-6; bf16x3 tensor-core convs y 1.9e-5 / dx 3.9e-3;
    fp16-operand PAM (round 1: single fp16 logits, bf16 gradient operands) y 8.5e-5 / dx 3.7e-2; 'fp16x3' (hi+lo split logits, fp16 gradient
    operands): see profiles/r02_precision_modes.json.  Why a 1e-4 output difference is a 1e-2 gradient difference in this network (ReLU sign
    flips: gradient error ~ sqrt(output error)) is measured on the CPU by tools/precision_bisect.py and asserted in tests/test_gpu_quantised.py."""
    from gan_danet_b200 import engine as E
    g = golden("generator_cin46_8x16")
    G = _make_generator(g["seed"], g["gamma"], precision)
    old = E.conv_precision
    E.set_conv_precision(conv)
    try:
        y, dx, grads = _fwd_bwd(G, g["x"], g["r"])
    finally:
        E.set_conv_precision(old)
    assert rel_err(y, g["y"]) < tol_y, rel_err(y, g["y"])
    assert rel_err(dx, g["dx"]) < tol_g, rel_err(dx, g["dx"])
    exact = conv == "fp32" and precision == "fp32"
    bad = []
    for k, ref_norm in g["grad_norms"].items():
        # d/d(key.bias) is analytically zero; the attention gammas are sums of ~1e5 cancelling terms (the reference's own
        # fp32 run is 3.5e-2 off its fp64 run on such tensors): asserted only for the fp32 engine
        if "key.bias" in k or (not exact and k.endswith("attention.gamma")):
            continue
        got = float(grads[k].double().norm())
        if abs(got - ref_norm) > 5 * tol_g * max(ref_norm, 1e-9):
            bad.append((k, got, ref_norm))
    assert not bad, bad[:5]
    small = {k: v for k, v in g["grads_small"].items() if "key.bias" not in k and (exact or not k.endswith("attention.gamma"))}
    num = sum(float((grads[k].double().cpu() - v.double()).norm() ** 2) for k, v in small.items())
    den = sum(float(v.double().norm() ** 2) for v in small.values())
    assert (num / den) ** 0.5 < tol_g, (num / den) ** 0.5          # whole-vector error over the stored gradient tensors
    worst = max((rel_err(grads[k], v), k) for k, v in small.items())
    # per-tensor errors of cancellation-dominated tensors (biases that BN removes) are large in the reference too
    assert worst[0] < 20 * tol_g, worst
    sd = G.state_dict()
    for k, v in g["buffers_after"].items():
        assert rel_err(sd[k], v) < 1e-4, k
    assert int(sd["initial.1.num_batches_tracked"]) == 1


def test_generator_deterministic_and_eval(golden):
    g = golden("generator_cin46_8x16")
    G = _make_generator(g["seed"], g["gamma"], "fp16")
    y1, dx1, gr1 = _fwd_bwd(G, g["x"], g["r"])
    G2 = _make_generator(g["seed"], g["gamma"], "fp16")
    y2, dx2, gr2 = _fwd_bwd(G2, g["x"], g["r"])
    assert torch.equal(y1, y2) and torch.equal(dx1, dx2)
    assert all(torch.equal(gr1[k], gr2[k]) for k in gr1)
    G.eval()
    with torch.no_grad():
        ye = G(g["x"].to(DEV))
    assert ye.shape == (2, 1, 32, 64) and torch.isfinite(ye).all()


def test_discriminator(golden):
    import gan_danet_b200 as P
    g = golden("discriminator_64x128")
    torch.manual_seed(g["seed"])
    D = P.Discriminator1()
    for mod in (D.conv1, D.conv2, D.conv3, D.conv4, D.fc2):
        mod.apply(P.weights_init_normal)
    D._materialise_fc1(g["x"])
    D = D.to(DEV)
    x = g["x"].to(DEV).requires_grad_(True)
    z = D(x)
    z.backward(torch.tensor([[1.0], [-0.5]], device=DEV))
    assert rel_err(z, g["logits"]) < 1e-4, rel_err(z, g["logits"])
    assert rel_err(x.grad, g["dx"]) < 1e-3
    for k, p in D.named_parameters():
        assert abs(float(p.grad.double().norm()) - g["grad_norms"][k]) < 2e-3 * g["grad_norms"][k], k
    for k, v in g["grads_small"].items():
        assert rel_err(dict(D.named_parameters())[k].grad, v) < 2e-3, k


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("fp16", 1e-2), ("fp16x3", 1e-2)])
def test_train_two_steps(golden, precision, tol):
    """Two full G+D steps (GAN_DANet_train.ipynb:225-269) against the reference modules + torch.optim.AdamW in float64.
    north_star bar: losses within 1 %."""
    import gan_danet_b200 as P
    from gan_danet_b200.synthetic import make_batch
    from gan_danet_b200.trainer import GANTrainer
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
    G, D = G.to(DEV), D.to(DEV)
    perc.vgg.to(DEV)
    perc.device = torch.device(DEV)
    G.set_pam_precision(precision)
    tr = GANTrainer(G, D, perc, epochs=g["epochs"], eval_ssim=True)
    tr.epoch = g["epoch"]
    for step in range(2):
        out = tr.train_step(lr05.to(DEV), real.to(DEV), aux.to(DEV))
        ref = g["history"][step]
        for k in ("loss_D", "loss_G", "adv", "pixel", "ssim", "tv", "perceptual"):
            got = float(out[k])
            assert abs(got - ref[k]) <= tol * max(abs(ref[k]), 1e-3), (step, k, got, ref[k])
    assert rel_err(G.final.weight, g["final_w"]) < 10 * tol
    assert rel_err(D.fc2.weight, g["d_fc2_w"]) < 10 * tol
    assert rel_err(G.initial[1].running_mean, g["initial_bn_rm"]) < 1e-4


def test_train_trajectory_free_running(oracle):
    """north_star: "loss trajectory within 1 %".  Free-running G+D training (GAN_DANet_train.ipynb:225-269) on four different batches
    per "epoch", fp32 engine, against the float64 CPU oracle started from the same weights.  SURVEY 8(c) addendum: even the
    reference run in fp32 against itself in fp64 stays within 1 % only for about ten steps (Discriminator1 has no normalisation
    and AdamW amplifies round-off; measured here: 5e-4 at step 4, 0.5 % or 1.6 % at step 7 depending on the order in which autograd sums dL/dhr), so the horizon asserted is 5 steps
    at 1 % on both losses (5 % at the sixth), post-step weights at 5e-3."""
    import gan_danet_b200 as P
    from gan_danet_b200 import engine as E
    from gan_danet_b200.synthetic import make_batch
    from gan_danet_b200.trainer import GANTrainer
    steps, epochs = 6, 150
    batches = [make_batch(10 * i, 2, 8, 16) for i in range(4)]
    torch.manual_seed(11)
    G = P.FlexibleUpsamplingModule(46)
    D = P.Discriminator1()
    G.apply(P.weights_init_normal)
    for mod in (D.conv1, D.conv2, D.conv3, D.conv4, D.fc2):
        mod.apply(P.weights_init_normal)
    D._materialise_fc1(batches[0][1])
    with torch.no_grad():
        for n, p in G.named_parameters():
            if n.endswith("gamma"):
                p.fill_(0.05)              # attention path live
    torch.manual_seed(12)
    perc = P.PerceptualLoss(pretrained=False, device=torch.device("cpu"))
    d64 = lambda sd: {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in sd.items()}  # noqa: E731
    st = oracle.TrainState(d64(G.state_dict()), d64(D.state_dict()), {k: v.double() for k, v in perc.vgg.state_dict().items()})
    ref = [oracle.train_step(st, *(t.double() for t in batches[i % 4]), epoch=3, epochs=epochs) for i in range(steps)]
    old = E.conv_precision
    E.set_conv_precision("fp32")
    try:
        G, D = G.to(DEV), D.to(DEV)
        perc.vgg.to(DEV)
        perc.device = torch.device(DEV)
        G.set_pam_precision("fp32")
        tr = GANTrainer(G, D, perc, epochs=epochs)
        tr.epoch = 3
        for i in range(steps):
            out = tr.train_step(*(t.to(DEV) for t in batches[i % 4]))
            for k in ("loss_D", "loss_G"):
                got = float(out[k])
                # measured realisations of this chaotic run (tools/trajectory_probe.py): 5e-8, 4e-6, 5e-5, ., 4e-4 for steps 0-4, then 0.5-2.3 % at
                # steps 5-7 depending on float32 summation orders that are irrelevant per step (the teacher-forced 200-step test holds 5e-6)
                tol = 1e-2 if i < 5 else 5e-2
                assert abs(got - ref[i][k]) <= tol * max(abs(ref[i][k]), 1e-3), (i, k, got, ref[i][k])
        torch.cuda.synchronize()
    finally:
        E.set_conv_precision(old)
    assert rel_err(G.final.weight, st.g["final.weight"]) < 5e-3
    assert rel_err(D.fc2.weight, st.d["fc2.weight"]) < 5e-3


def test_pam_properties_full_size():
    """Size-independent properties of the fused kernel at the BASELINE grid (N = 8192, C = 184):
    with a constant value map softmax rows sum to one, so gamma*O + x == gamma*v + x exactly up to fp16 rounding;
    and the output is linear in V."""
    from gan_danet_b200 import engine as E
    from gan_danet_b200._lib import PREC_FP16
    B, H, W, C, d = 2, 64, 128, 184, 23
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(B, H, W, C, generator=gen).to(DEV)
    q = (2.0 * torch.randn(B, H, W, d, generator=gen)).to(DEV)
    k = (2.0 * torch.randn(B, H, W, d, generator=gen)).to(DEV)
    gamma = torch.full((1,), 0.5, device=DEV)

    def run(v):
        t = E.Tape(record=False)
        return E.op_pam_core(t, E.Var(x), E.Var(q), E.Var(k), E.Var(v), E.Var(gamma), precision=PREC_FP16).t

    ones = torch.full((B, H, W, C), 0.75, device=DEV)
    y = run(ones)
    assert float((y - (x + 0.5 * 0.75)).abs().max()) < 2e-3
    v1 = torch.randn(B, H, W, C, generator=gen).to(DEV)
    v2 = torch.randn(B, H, W, C, generator=gen).to(DEV)
    y1, y2, y12 = run(v1) - x, run(v2) - x, run(v1 + v2) - x
    assert rel_err(y12, y1 + y2) < 3e-3


@pytest.mark.parametrize("B,H,W,C", [(2, 9, 13, 8), (1, 16, 32, 64), (1, 3, 4, 4), (2, 4, 6, 4), (1, 2, 2, 8), (1, 6, 2, 4), (1, 29, 7, 8), (2, 40, 16, 4)])
def test_resample_vector_paths(B, H, W, C):
    """128-bit kernels (C % 4 == 0) for nn.Upsample(2,'bicubic') (generator.py:221,225) and MaxPool2d(2,2) (VGG19 of
    losses.py:58), forward and backward, against ATen on the same device in float64 (same formulas as the CPU oracle)."""
    import torch.nn.functional as F
    from gan_danet_b200 import engine as E
    g = torch.Generator().manual_seed(H * 100 + W)
    x = torch.randn(B, H, W, C, generator=g).to(DEV)
    tape = E.Tape()
    xv = E.Var(x)
    up = E.op_bicubic_up2(tape, xv)
    r = torch.randn(up.t.shape, generator=g).to(DEV)
    up.g = r
    tape.backward()
    xd = x.double().permute(0, 3, 1, 2).requires_grad_(True)
    ref = F.interpolate(xd, scale_factor=2, mode="bicubic", align_corners=False)
    ref.backward(r.double().permute(0, 3, 1, 2))
    assert rel_err(up.t.permute(0, 3, 1, 2), ref) < 1e-6
    assert rel_err(xv.g.permute(0, 3, 1, 2), xd.grad) < 1e-6
    tape = E.Tape()
    xv = E.Var(x)
    mp = E.op_maxpool2(tape, xv)
    r2 = torch.randn(mp.t.shape, generator=g).to(DEV)
    mp.g = r2
    tape.backward()
    xd = x.double().permute(0, 3, 1, 2).requires_grad_(True)
    ref = F.max_pool2d(xd, 2, 2)
    ref.backward(r2.double().permute(0, 3, 1, 2))
    assert torch.equal(mp.t.permute(0, 3, 1, 2).double(), ref.detach())
    assert rel_err(xv.g.permute(0, 3, 1, 2), xd.grad) < 1e-7


def test_srgand(golden):
    """SRGAND (models/discriminator.py:8-54): 4x4 stride-2 convolutions, BatchNorm + LeakyReLU(0.2), residual add, global average pool, Linear --
    forward logits, input gradient, parameter gradients and BN running statistics against the reference's float64 run (fp32 engine: the 4x4
    kernels are outside the tensor-core tiling).  This is the discriminator with the BatchNorm + LeakyReLU epilogue the north star names."""
    import gan_danet_b200 as P
    g = golden("srgand_dim8_128x128")
    torch.manual_seed(g["seed"])
    D = P.SRGAND(dim=g["dim"], in_channels=1)
    D.apply(P.weights_init_normal)
    assert list(D.state_dict().keys()) == g["keys"]
    D = D.to(DEV).train()
    x = g["x"].to(DEV).requires_grad_(True)
    z = D(x)
    z.backward(torch.tensor([[1.0], [-0.5]], device=DEV))
    torch.cuda.synchronize()
    assert rel_err(z, g["logits"]) < 1e-4, rel_err(z, g["logits"])
    assert rel_err(x.grad, g["dx"]) < 2e-3, rel_err(x.grad, g["dx"])
    params = dict(D.named_parameters())
    bad = [(k, float(params[k].grad.double().norm()), n) for k, n in g["grad_norms"].items()
           if abs(float(params[k].grad.double().norm()) - n) > 5e-3 * max(n, 1e-7) and not k.endswith(".bias")]
    assert not bad, bad[:5]
    # conv biases in front of a train-mode BatchNorm have an analytically zero gradient (the batch mean removes them): noise in the reference too
    for k, v in g["grads_small"].items():
        if k.startswith("conv") and k.endswith(".bias") and k not in ("conv1.bias",):
            assert float(params[k].grad.abs().max()) < 1e-3 * max(float(params[k.replace(".bias", ".weight")].grad.abs().max()), 1e-12), k
        else:
            assert rel_err(params[k].grad, v) < 5e-3, (k, rel_err(params[k].grad, v))
    sd = D.state_dict()
    for k, v in g["buffers_after"].items():
        assert rel_err(sd[k], v) < 1e-4, (k, rel_err(sd[k], v))


@pytest.mark.parametrize("with_skip", [True, False])
def test_generator_head_tap_planes(with_skip):
    """engine.op_upsample_skip_final (final 3x3 conv's channel reduction hoisted in front of the bicubic x2 / bilinear skip resize: no 64-channel
    full-resolution tensor) against the unfused chain up2 -> (+ resize(s)) -> conv3x3 on the same kernels, and against ATen in float64: output and
    all gradients.  Exact in real arithmetic; asserted at 2e-5 (another fp32 summation order)."""
    import torch.nn.functional as F
    from gan_danet_b200 import engine as E
    g = torch.Generator().manual_seed(31)
    B, H, W, C = 2, 12, 20, 64
    u = torch.randn(B, H, W, C, generator=g).to(DEV)
    s = torch.randn(B, H // 2, W // 2, C, generator=g).to(DEV) if with_skip else None
    w = (0.1 * torch.randn(1, C, 3, 3, generator=g)).to(DEV)
    b = torch.randn(1, generator=g).to(DEV)
    r = torch.randn(B, 2 * H, 2 * W, 1, generator=g).to(DEV)
    tape = E.Tape()
    uv, sv, wv, bv = E.Var(u), (E.Var(s) if with_skip else None), E.Var(w), E.Var(b)
    assert E.head_tap_planes_ok(uv, sv, wv)
    y = E.op_upsample_skip_final(tape, uv, sv, wv, bv)
    y.g = r.clone()
    tape.backward()
    torch.cuda.synchronize()
    ud, wd, bd = (t.detach().cpu().double().requires_grad_(True) for t in (u, w, b))
    x2 = F.interpolate(ud.permute(0, 3, 1, 2), scale_factor=2, mode="bicubic", align_corners=False)
    sd = None
    if with_skip:
        sd = s.detach().cpu().double().requires_grad_(True)
        x2 = x2 + F.interpolate(sd.permute(0, 3, 1, 2), size=(2 * H, 2 * W), mode="bilinear", align_corners=False)
    ref = F.conv2d(x2, wd, bd, padding=1).permute(0, 2, 3, 1)
    (ref * r.cpu().double()).sum().backward()
    assert rel_err(y.t, ref) < 2e-5, rel_err(y.t, ref)
    assert rel_err(uv.g, ud.grad) < 2e-5 and rel_err(wv.g, wd.grad) < 2e-5 and rel_err(bv.g, bd.grad) < 2e-5
    if with_skip:
        assert rel_err(sv.g, sd.grad) < 2e-5
