"""relu1_1 of the perceptual loss evaluated from the two single-channel images instead of stored maps (gdn_thin_conv_tap_pair, gdn_thin_conv_tap_dgrad;
models/losses.py::_tap1_pair, _frozen_conv1_tap).  Reference: models/losses.py:58-72 with conv1_1 applied to x.repeat(1, 3, 1, 1) (= the channel-summed
weight on one channel, losses.py:64-65) and feature_layers containing 1."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.mark.parametrize("shape", [(2, 32, 64), (1, 45, 24), (3, 16, 32), (1, 256, 512)], ids=lambda s: "x".join(map(str, s)))
def test_tap_kernels_vs_float64(shape):
    """The pair kernel and the mask-driven data gradient against float64 autograd of  mean|relu(conv a) - relu(conv b)| + <dy, relu(conv a)>."""
    from gan_danet_b200 import _lib as L, engine as E
    B, H, W = shape
    C = 64
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(B * H + W)
    a = torch.randn(B, H, W, generator=g).to(dev)
    b = (a.cpu() + 0.3 * torch.randn(B, H, W, generator=g)).to(dev)
    w = (0.4 * torch.randn(C, 1, 3, 3, generator=g)).to(dev)
    bias = (0.1 * torch.randn(C, generator=g)).to(dev)
    dy = (1e-6 * torch.randn(B, H, W, C, generator=g)).to(dev)
    lib = E._lib(a)
    loss = torch.zeros(1, device=dev)
    ws = E.dot_ws(dev)
    y16a = torch.empty(B * H * W, C, dtype=torch.bfloat16, device=dev)
    y16b = torch.empty_like(y16a)
    mask = torch.empty(lib.gdn_thin_conv_tap_mask_bytes(B, H, W, C), dtype=torch.uint8, device=dev)
    L.check(lib.gdn_thin_conv_tap_pair(a.data_ptr(), b.data_ptr(), w.data_ptr(), bias.data_ptr(), B, H, W, C, loss.data_ptr(), 1, 1.0, y16a.data_ptr(), y16b.data_ptr(),
                                       mask.data_ptr(), ws.data_ptr(), ws.numel(), E._stream()), "tap_pair")
    gx = torch.empty(B, H, W, device=dev)
    n = B * H * W * C
    L.check(lib.gdn_thin_conv_tap_dgrad(dy.data_ptr(), C, mask.data_ptr(), w.data_ptr(), 1.0 / n, gx.data_ptr(), None, B, H, W, C, E._stream()), "tap_dgrad")
    # the bf16 operands of conv1_2 == the bf16 copies the stored-map path writes next to the fp32 maps
    ref16 = []
    y32 = torch.empty(B, H, W, C, device=dev)
    for img in (a, b):
        r16 = torch.empty_like(y16a)
        L.check(lib.gdn_thin_conv_expand_p(img.data_ptr(), w.data_ptr(), bias.data_ptr(), y32.data_ptr(), C, None, 0, B, H, W, C, H, W, 1, 1, 0, 1, 0.0, r16.data_ptr(), E._stream()), "expand")
        ref16.append(r16)
    torch.cuda.synchronize()
    assert torch.equal(y16a, ref16[0]) and torch.equal(y16b, ref16[1])
    ad = a.double().unsqueeze(1).requires_grad_(True)
    fa = F.relu(F.conv2d(ad, w.double(), bias.double(), padding=1))
    fb = F.relu(F.conv2d(b.double().unsqueeze(1), w.double(), bias.double(), padding=1))
    l1 = (fa - fb).abs().mean()
    tot = l1 + (fa * dy.double().permute(0, 3, 1, 2)).sum()
    tot.backward()
    assert abs(float(loss) - float(l1)) <= 2e-6 * float(l1)
    # gate / sign flips of fp32 vs float64 features touch isolated elements only
    assert rel(gx, ad.grad[:, 0]) < 2e-4
    # the 2-bit codes against the float64 maps (isolated borderline elements may differ)
    codes = torch.stack([(mask.view(B, H, W, C // 4) >> (2 * e)) & 3 for e in range(4)], dim=-1).reshape(B, H, W, C).long()
    fa_, fb_ = fa.detach().permute(0, 2, 3, 1), fb.permute(0, 2, 3, 1)
    want = torch.where(fa_ > 0, 2 + torch.sign(fa_ - fb_).long(), torch.zeros_like(codes))
    assert float((codes != want).double().mean()) < 1e-4


@pytest.mark.parametrize("shape", [(2, 32, 64), (1, 48, 64)], ids=lambda s: "x".join(map(str, s)))
def test_perceptual_recompute_matches_stored(shape):
    """PerceptualLoss in the bf16 product mode with and without the recomputation: same loss (summation order only), same input gradient."""
    import gan_danet_b200 as P
    from gan_danet_b200 import engine as E
    B, H, W = shape
    dev = torch.device("cuda", 0)
    old_prec, old_flag = E.conv_precision, E.vgg_tap1_recompute
    E.set_conv_precision("bf16")
    try:
        torch.manual_seed(5)
        crit = P.PerceptualLoss(feature_layers=(1, 6, 11, 20), pretrained=False, device=dev)
        g = torch.Generator().manual_seed(11)
        x0 = torch.randn(B, 1, H, W, generator=g).to(dev)
        y = (x0.cpu() + 0.2 * torch.randn(B, 1, H, W, generator=g)).to(dev)
        res = {}
        for flag in (False, True):
            E.vgg_tap1_recompute = flag
            x = x0.clone().requires_grad_(True)
            assert crit._tap1_recompute_ok(x) == flag
            loss = crit(x, y)
            loss.backward()
            torch.cuda.synchronize()
            res[flag] = (float(loss), x.grad.clone())
        assert abs(res[True][0] - res[False][0]) <= 2e-6 * abs(res[False][0])
        assert rel(res[True][1], res[False][1]) < 1e-6
    finally:
        E.set_conv_precision(old_prec)
        E.vgg_tap1_recompute = old_flag


def test_perceptual_recompute_falls_back_on_ineligible_grid():
    """A grid whose deeper VGG levels leave the bf16 feature-map path (odd width after three poolings) keeps the stored relu1_1 map."""
    import gan_danet_b200 as P
    from gan_danet_b200 import engine as E
    dev = torch.device("cuda", 0)
    old_prec = E.conv_precision
    E.set_conv_precision("bf16")
    try:
        torch.manual_seed(5)
        crit = P.PerceptualLoss(feature_layers=(1, 6, 11, 20), pretrained=False, device=dev)
        x = torch.randn(1, 1, 48, 40, device=dev, requires_grad=True)
        assert not crit._tap1_recompute_ok(x)
        crit(x, torch.randn(1, 1, 48, 40, device=dev)).backward()
        assert torch.isfinite(x.grad).all()
    finally:
        E.set_conv_precision(old_prec)
