"""tcgen05 implicit-GEMM convolution (conv_tc.cu) against a float64 convolution of the same operands.

``bf16``  : the kernel rounds x, w (and dy) to bf16 and accumulates in fp32, so it must agree with the float64 result computed
            from the SAME bf16-rounded operands to fp32-accumulation accuracy (1e-5), and with the unrounded one to ~1e-2.
``bf16x3``: hi+lo split operands; must agree with the float64 result of the UNROUNDED operands to 5e-5.
Shapes: the reference's layer shapes (generator.py:34,63,148,188; discriminator.py:63-65) on small and ragged grids
(45x22 is the authors' real grid).
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


CASES = [
    # B, Cin, Cout, H, W, k, stride, pad
    (2, 46, 64, 16, 32, 3, 1, 1),
    (1, 88, 24, 45, 22, 3, 1, 1),
    (2, 320, 160, 8, 16, 3, 1, 1),
    (2, 160, 80, 8, 16, 1, 1, 0),
    (3, 184, 64, 64, 128, 1, 1, 0),
    (2, 64, 128, 32, 64, 3, 2, 1),
    (1, 128, 256, 45, 22, 3, 2, 1),
    (2, 256, 512, 16, 32, 3, 2, 1),
    (2, 64, 1, 32, 64, 3, 1, 1),
    (1, 64, 64, 128, 256, 3, 1, 1),
    (2, 1, 64, 64, 128, 3, 2, 1),
    (1, 136, 24, 16, 128, 3, 1, 1),   # halo-reuse variant of the forward kernel (one-row tiles of 128 pixels, narrow N)
    (2, 64, 1, 8, 256, 3, 1, 1),
    (1, 88, 24, 5, 200, 3, 1, 1),     # Discriminator1.conv1: forward on the CUDA cores, gradients on tensor cores
    (1, 352, 176, 4, 128, 3, 1, 1),   # data gradient 176 -> 352: halo variant with a wide output tile (n_tile 176, three channel chunks)
    (1, 112, 24, 3, 256, 3, 1, 1),    # data gradient 24 -> 112: halo variant with one channel chunk
]


@pytest.mark.parametrize("precision", ["bf16", "bf16x3"])
@pytest.mark.parametrize("case", CASES, ids=[f"b{c[0]}_{c[1]}to{c[2]}_{c[3]}x{c[4]}_k{c[5]}s{c[6]}" for c in CASES])
def test_conv_tc(case, precision):
    from gan_danet_b200 import engine as E
    from gan_danet_b200._lib import ACT_LRELU, ACT_NONE
    B, Cin, Cout, H, W, k, stride, pad = case
    dev = torch.device("cuda", 0)
    g = torch.Generator(device="cpu").manual_seed(hash(case) & 0xFFFF)
    x = torch.randn(B, H, W, Cin, generator=g).to(dev)
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).to(dev)
    bias = torch.randn(Cout, generator=g).to(dev)
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    res = torch.randn(B, Ho, Wo, Cout, generator=g).to(dev)
    dy = torch.randn(B, Ho, Wo, Cout, generator=g).to(dev)
    old = E.conv_precision
    E.set_conv_precision(precision)
    try:
        thin = Cin == 1 or Cout == 1          # 1 <-> C channel convs run on the fp32 thin-conv kernels (thin_conv.cu) in every precision mode
        fwd_tc = E.tc_eligible(Cin, Cout, k, k, stride, Ho, Wo) and not thin
        y = torch.empty(B, Ho, Wo, Cout, device=dev)
        ctx = E.conv_forward(x, w, y, stride=stride, pad=pad, bias=bias, act=ACT_LRELU, slope=0.2, res=res)
        assert ctx.tc == (fwd_tc or Cout == 1)      # the C -> 1 thin kernel has no activation epilogue: that call takes the tensor-core path
        y2 = torch.empty(B, Ho, Wo, Cout, device=dev)
        ctx2 = E.conv_forward(x, w, y2, stride=stride, pad=pad)
        gw = torch.empty_like(w)
        gx = torch.randn(B, H, W, Cin, generator=g).to(dev)
        gx0 = gx.clone()
        E.conv_backward(ctx2, dy, x, w, stride=stride, pad=pad, gw=gw, gx=gx, gx_accumulate=True)
        torch.cuda.synchronize()
    finally:
        E.set_conv_precision(old)

    rnd = bf16_round if (precision == "bf16" and not thin) else (lambda t: t)
    xr, wr, dyr = rnd(x).double(), rnd(w).double(), rnd(dy).double()
    xn = xr.permute(0, 3, 1, 2).requires_grad_(True)
    wn = wr.clone().requires_grad_(True)
    yref = F.conv2d(xn, wn, None, stride=stride, padding=pad)
    yref.backward(dyr.permute(0, 3, 1, 2))
    y_plain = yref.detach().permute(0, 2, 3, 1)
    y_full = F.leaky_relu(y_plain + bias.double(), 0.2) + res.double()
    tol = 2e-5 if precision == "bf16" else 5e-5
    ftol = tol
    assert rel(y2, y_plain) < ftol, ("fwd", rel(y2, y_plain))
    assert rel(y, y_full) < (1e-2 if Cout == 1 else ftol), ("fwd+epilogue", rel(y, y_full))
    assert rel(gw, wn.grad) < tol, ("wgrad", rel(gw, wn.grad))
    wtol = tol
    gx_ref = xn.grad.permute(0, 2, 3, 1) + gx0.double()
    assert rel(gx, gx_ref) < wtol, ("dgrad", rel(gx, gx_ref))   # Cout < 16: data gradient on the fp32 engine as well


def test_perceptual_bf16_feature_storage():
    """Frozen VGG19 branch (losses.py:58-73) in conv precision 'bf16': keeping the untapped feature maps in bf16 only (conv epilogue
    writes the next conv's packed operand, bf16 max-pool, ReLU backward gated by the bf16 map) must reproduce the fp32-storage path:
    both round the same fp32 values to bf16 at the same places, so loss and input gradient agree to fp32 round-off."""
    import gan_danet_b200 as P
    from gan_danet_b200 import engine as E
    dev = torch.device("cuda", 0)
    torch.manual_seed(3)
    perc = P.PerceptualLoss(pretrained=False, device=torch.device("cpu"))
    perc.vgg.to(dev)
    perc.device = dev
    g = torch.Generator().manual_seed(5)
    x0 = torch.randn(2, 1, 64, 128, generator=g).to(dev)
    y = torch.randn(2, 1, 64, 128, generator=g).to(dev)
    old_prec, old_flag = E.conv_precision, E.bf16_feature_storage
    res = []
    try:
        E.set_conv_precision("bf16")
        for flag in (False, True):
            E.bf16_feature_storage = flag
            x = x0.clone().requires_grad_(True)
            loss = perc(x, y)
            loss.backward()
            torch.cuda.synchronize()
            res.append((float(loss), x.grad.detach().clone()))
    finally:
        E.set_conv_precision(old_prec)
        E.bf16_feature_storage = old_flag
    (l0, g0), (l1, g1) = res
    assert abs(l1 - l0) <= 1e-5 * abs(l0), (l0, l1)
    assert rel(g1, g0) < 1e-5, rel(g1, g0)


@pytest.mark.parametrize("precision", ["bf16", "bf16x3"])
@pytest.mark.parametrize("train", [True, False])
def test_bn_relu_fused_into_operand_packing(precision, train):
    """DenseBlock (generator.py:29-54) and TransitionLayer (:57-67) with BatchNorm + ReLU applied while the convolution's bf16
    operand is packed (engine.op_bn_act_conv: the normalised activation never exists in fp32) against the unfused launches
    (BN + ReLU kernel, then packing): the same fused multiply-add and the same rounding, so outputs, input gradient, every
    parameter gradient and the running statistics must agree to float32 round-off."""
    from gan_danet_b200 import engine as E
    from gan_danet_b200.models import generator as PG
    import gan_danet_b200 as P
    dev = "cuda:0"
    gen = torch.Generator().manual_seed(21)
    x = torch.randn(2, 64, 16, 24, generator=gen)
    r1 = torch.randn(2, 160, 16, 24, generator=gen)

    def run(fuse):
        torch.manual_seed(5)
        blk, tr = PG.DenseBlock(4, 64, 24), PG.TransitionLayer(160, 80)
        for m in (blk, tr):
            m.apply(P.weights_init_normal)
        seq = torch.nn.Sequential(blk, tr).to(dev)
        seq.train(train)
        old, oldf = E.conv_precision, E.fuse_bn_into_pack
        E.set_conv_precision(precision)
        E.fuse_bn_into_pack = fuse
        try:
            xg = x.to(dev).requires_grad_(train)
            if train:
                y = seq(xg)
                y.backward(torch.ones_like(y) * 0.5 + y.detach() * 0.1)
            else:
                with torch.no_grad():
                    y = seq(xg)
            torch.cuda.synchronize()
        finally:
            E.set_conv_precision(old)
            E.fuse_bn_into_pack = oldf
        grads = {k: p.grad.detach() for k, p in seq.named_parameters()} if train else {}
        bufs = {k: v.detach().clone() for k, v in seq.state_dict().items() if "running" in k}
        return y.detach(), (xg.grad.detach() if train else None), grads, bufs

    yf, dxf, gf, bf = run(True)
    yu, dxu, gu, bu = run(False)
    assert rel(yf, yu) < 1e-6, rel(yf, yu)
    for k in bu:
        assert rel(bf[k], bu[k]) < 1e-6, k
    if train:
        assert rel(dxf, dxu) < 1e-5, rel(dxf, dxu)
        for k in gu:
            assert rel(gf[k], gu[k]) < 1e-5 or float(gu[k].abs().max()) < 1e-6, (k, rel(gf[k], gu[k]))


def test_conv_bn_relu_packed_gradient():
    """conv -> BN -> ReLU chains of the generator (initial, DANet fuse, upsample: generator.py:147-150,187-190,218-224) in the bf16 product mode:
    BatchNorm's backward hands the convolution its dz as a bf16 operand only (gdn_bn_bwd_apply16) instead of an fp32 tensor that is packed
    afterwards -- the same round-to-nearest of the same fp32 values, so every gradient is bitwise that of the unfused path."""
    from gan_danet_b200 import engine as E
    import gan_danet_b200 as P
    dev = "cuda:0"
    from gan_danet_b200.synthetic import fast_batch
    lr05, _, aux = fast_batch(3, 2, 16, 32)

    def run(flag):
        torch.manual_seed(0)
        G = P.FlexibleUpsamplingModule(46)
        G.apply(P.weights_init_normal)
        with torch.no_grad():
            for n, p in G.named_parameters():
                if n.endswith("gamma"):
                    p.fill_(0.05)
        G = G.to(dev).train()
        G.set_pam_precision("fp16")
        old, oldf = E.conv_precision, E.conv_bn_packed_grad
        E.set_conv_precision("bf16")
        E.conv_bn_packed_grad = flag
        try:
            import sys, os
            sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
            import gan_danet_oracle as oracle
            x = oracle.prepare_input(lr05, aux).to(dev).requires_grad_(True)
            y = G(x)
            y.backward(torch.ones_like(y) * 0.25 + 0.1 * y.detach())
            torch.cuda.synchronize()
        finally:
            E.set_conv_precision(old)
            E.conv_bn_packed_grad = oldf
        return y.detach(), x.grad.detach(), {k: p.grad.detach() for k, p in G.named_parameters()}

    y1, dx1, g1 = run(True)
    y0, dx0, g0 = run(False)
    assert torch.equal(y1, y0) and torch.equal(dx1, dx0)
    bad = [k for k in g0 if not torch.equal(g1[k], g0[k])]
    assert not bad, bad[:5]


@pytest.mark.parametrize("precision", ["bf16", "bf16x3"])
@pytest.mark.parametrize("B,Cin,Cout,H,W,k", [(2, 136, 24, 64, 128, 3), (1, 64, 24, 45, 22, 3), (2, 160, 24, 16, 32, 3), (1, 88, 16, 9, 13, 1), (1, 152, 32, 8, 200, 3)])
def test_wgrad_role_swap(B, Cin, Cout, H, W, k, precision):
    """Weight gradient of the narrow (growth) convolutions of the dense blocks (generator.py:34): the role-swapped launch (input channels as the
    UMMA M dimension, taps mirrored) against the direct one on the same bf16 operands -- the same products, another accumulation order."""
    from gan_danet_b200 import _lib, engine as E
    dev = torch.device("cuda", 0)
    g = torch.Generator(device="cpu").manual_seed(B * 100 + Cin)
    x = torch.randn(B, H, W, Cin, generator=g).to(dev)
    dy = torch.randn(B, H, W, Cout, generator=g).to(dev)
    lib = _lib.lib_for_device(0)
    old = E.conv_precision
    E.set_conv_precision(precision)
    try:
        out = []
        for swap in (1, 0):
            prev = lib.gdn_conv_tc_set_wgrad_swap(swap)
            try:
                gw = torch.empty(Cout, Cin, k, k, device=dev)
                E.wgrad_tc_raw(E.pack_act(dy), E.pack_act(x), gw, B=B, in_hw=(H, W), out_hw=(H, W), cin=Cin, cout=Cout, kh=k, kw=k, stride=1, pad=k // 2)
                torch.cuda.synchronize()
                out.append(gw)
            finally:
                lib.gdn_conv_tc_set_wgrad_swap(prev)
        err = float((out[0].double() - out[1].double()).norm() / out[1].double().norm())
        assert err < 1e-5, err
        xr, dyr = (bf16_round(t) if precision == "bf16" else t for t in (x, dy))
        ref = torch.nn.grad.conv2d_weight(xr.double().permute(0, 3, 1, 2), (Cout, Cin, k, k), dyr.double().permute(0, 3, 1, 2), padding=k // 2)
        assert float((out[0].double() - ref).norm() / ref.norm()) < (1e-5 if precision == "bf16" else 1e-4)
    finally:
        E.set_conv_precision(old)


WGCOL_CASES = [(1, 8, 16, 64, 24), (2, 16, 128, 136, 24), (1, 45, 22, 88, 24), (1, 5, 200, 88, 24), (1, 32, 64, 80, 20), (1, 16, 128, 112, 8), (3, 64, 128, 160, 24), (1, 9, 13, 16, 3)]


@pytest.mark.parametrize("case", WGCOL_CASES, ids=[f"b{c[0]}_{c[1]}x{c[2]}_{c[3]}to{c[4]}" for c in WGCOL_CASES])
def test_wgrad_narrow_output_kernel(case):
    """conv_tc_wgrad_col_kernel (weight gradient of the DenseNet growth convolutions, generator.py:34: the shifted operand is dy, one MMA chain per pixel
    tile) against float64 on the same bf16-rounded operands (fp32-accumulation accuracy) and against the general weight-gradient kernel; bitwise
    repeatable; the accumulate flag adds."""
    from gan_danet_b200 import _lib, engine as E
    B, H, W, Cin, Cout = case
    dev = torch.device("cuda", 0)
    lib = _lib.lib_for_device(0)
    g = torch.Generator().manual_seed(Cin * 7 + H)
    x = torch.randn(B, H, W, Cin, generator=g).to(dev)
    dy = torch.randn(B, H, W, Cout, generator=g).to(dev)
    old = E.conv_precision
    E.set_conv_precision("bf16")
    try:
        xp, dyp = E.pack_act(x), E.pack_act(dy)
        res = {}
        for col, row in ((0, 1), (1, 1), (1, 1), (2, 0)):       # general kernel; narrow kernel (twice: determinism, accumulate); narrow kernel without the row variant
            prev, prev_row = lib.gdn_conv_tc_set_wgrad_col(1 if col else 0), lib.gdn_conv_tc_set_wgrad_col_row(row)
            try:
                gw = torch.zeros(Cout, Cin, 3, 3, device=dev)
                E.wgrad_tc_raw(dyp, xp, gw, B=B, in_hw=(H, W), out_hw=(H, W), cin=Cin, cout=Cout, kh=3, kw=3, pad=1)
                if col and col in res:
                    assert torch.equal(gw, res[col])        # deterministic
                    E.wgrad_tc_raw(dyp, xp, gw, B=B, in_hw=(H, W), out_hw=(H, W), cin=Cin, cout=Cout, kh=3, kw=3, pad=1, accumulate=True)
                    torch.testing.assert_close(gw, 2 * res[col], rtol=1e-6, atol=1e-6)
                else:
                    res[col] = gw.clone()
            finally:
                lib.gdn_conv_tc_set_wgrad_col(prev)
                lib.gdn_conv_tc_set_wgrad_col_row(prev_row)
        xr, dr = bf16_round(x).double().permute(0, 3, 1, 2), bf16_round(dy).double().permute(0, 3, 1, 2)
        w = torch.zeros(Cout, Cin, 3, 3, dtype=torch.float64, device=dev, requires_grad=True)
        (F.conv2d(xr, w, padding=1) * dr).sum().backward()
        assert rel(res[1], w.grad) < 1e-5 and rel(res[2], w.grad) < 1e-5
        assert rel(res[1], res[0]) < 2e-5
    finally:
        E.set_conv_precision(old)


COL_CASES = [(1, 16, 128, 136, 24), (2, 3, 256, 112, 24), (1, 8, 128, 64, 20), (1, 5, 128, 160, 8), (1, 7, 128, 248, 24), (3, 64, 128, 88, 24)]


@pytest.mark.parametrize("case", COL_CASES, ids=[f"b{c[0]}_{c[1]}x{c[2]}_{c[4]}to{c[3]}" for c in COL_CASES])
def test_dgrad_narrow_gradient_col_mode(case):
    """COL mode of the forward kernel (data gradient of the DenseNet growth convolutions, generator.py:34: one K = 3 x 80 product over the im2col'd
    24-channel gradient tile, weights resident) against float64 on the same bf16-rounded operands and against the per-tap kernel, with the
    accumulating epilogue (res = the gradient buffer itself) the dense block uses."""
    from gan_danet_b200 import _lib, engine as E
    B, H, W, Cin, Cout = case
    dev = torch.device("cuda", 0)
    lib = _lib.lib_for_device(0)
    g = torch.Generator().manual_seed(Cin * 3 + W)
    dy = torch.randn(B, H, W, Cout, generator=g).to(dev)
    w = (0.1 * torch.randn(Cout, Cin, 3, 3, generator=g)).to(dev)
    res0 = torch.randn(B, H, W, Cin, generator=g).to(dev)
    old = E.conv_precision
    E.set_conv_precision("bf16")
    try:
        dyp, wt = E.pack_act(dy), E.pack_weight(w, True)
        outs = []
        for col in (0, 1, 1):
            prev = lib.gdn_conv_tc_set_col(col)
            try:
                gx = res0.clone()
                E.conv_tc_raw(dyp, wt, gx, (H, W), cin=Cout, kh=3, kw=3, pad=1, transposed=True, res=gx)
                outs.append(gx)
            finally:
                lib.gdn_conv_tc_set_col(prev)
        ref = F.conv_transpose2d(bf16_round(dy).double().permute(0, 3, 1, 2), bf16_round(w).double(), padding=1).permute(0, 2, 3, 1) + res0.double()
        assert rel(outs[1], ref) < 1e-5 and rel(outs[0], ref) < 1e-5
        assert torch.equal(outs[1], outs[2])
    finally:
        E.set_conv_precision(old)
