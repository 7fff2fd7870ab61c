"""Single-launch statistics (gdn_bn_stats, gdn_bn_bwd_reduce_f, gdn_colsums_f: the reduction's last block finishes it) against the
multi-launch path they replace (gdn_colstats + gdn_bn_finalize, gdn_bn_bwd_reduce + gdn_sums_to_float) and against a float64 evaluation
of train-mode nn.BatchNorm2d statistics (reference generator.py:32,61,149,189,219,223).  Shapes: one block, one group, several groups,
the step's largest (262144 rows), ragged channel counts (scalar kernel) and channel slices of a wider buffer."""
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = [  # rows, C, pitch
    (40, 8, 8), (1000, 24, 24), (128 * 64, 64, 160), (45 * 22 * 3, 46, 46), (32 * 64 * 128, 184, 184), (70000, 17, 19), (4096, 1, 1),
]


def _coef(dev, C):
    return torch.empty((4, C), dtype=torch.float32, device=dev)


@pytest.mark.parametrize("case", CASES, ids=[f"m{m}_c{c}_p{p}" for m, c, p in CASES])
def test_bn_stats_single_launch(case):
    from gan_danet_b200 import _lib as L, engine as E
    M, C, pitch = case
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(M + C)
    buf = (torch.randn(M, pitch, generator=g) * 1.7 + 0.3).to(dev)
    x = buf[:, :C]
    w, b = torch.randn(C, generator=g).to(dev), torch.randn(C, generator=g).to(dev)
    lib = E._lib(x)
    rm0, rv0 = torch.randn(C, generator=g).to(dev), (torch.rand(C, generator=g) + 0.5).to(dev)
    # multi-launch path
    rm_a, rv_a, ca = rm0.clone(), rv0.clone(), _coef(dev, C)
    sums = E.colstats(x)
    L.check(lib.gdn_bn_finalize(sums.data_ptr(), M, C, w.data_ptr(), b.data_ptr(), 1e-5, 0.1, rm_a.data_ptr(), rv_a.data_ptr(),
                                ca[0].data_ptr(), ca[1].data_ptr(), ca[2].data_ptr(), ca[3].data_ptr(), E._stream()), "gdn_bn_finalize")
    # single launch, three times in a row (the counters must come back to zero; results bitwise repeatable)
    outs = []
    nbt = torch.zeros((), dtype=torch.int64, device=dev)
    for _ in range(3):
        rm_b, rv_b, cb = rm0.clone(), rv0.clone(), _coef(dev, C)
        ws = E.workspace("stat", lib.gdn_stat_fused_ws_bytes(M, C), dev)
        L.check(lib.gdn_bn_stats(x.data_ptr(), pitch, 0, M, C, w.data_ptr(), b.data_ptr(), 1e-5, 0.1, rm_b.data_ptr(), rv_b.data_ptr(), nbt.data_ptr(),
                                 cb[0].data_ptr(), cb[1].data_ptr(), cb[2].data_ptr(), cb[3].data_ptr(), ws.data_ptr(), E.stat_counters(dev).data_ptr(),
                                 E._stream()), "gdn_bn_stats")
        outs.append((rm_b, rv_b, cb))
    torch.cuda.synchronize()
    assert int(nbt) == 3
    assert int(E.stat_counters(dev).abs().sum()) == 0
    for rm_b, rv_b, cb in outs[1:]:
        assert torch.equal(cb, outs[0][2]) and torch.equal(rm_b, outs[0][0]) and torch.equal(rv_b, outs[0][1])
    rm_b, rv_b, cb = outs[0]
    # both paths accumulate in double: they agree to float rounding of the results
    torch.testing.assert_close(cb, ca, rtol=2e-6, atol=2e-6)
    torch.testing.assert_close(rm_b, rm_a, rtol=2e-6, atol=2e-6)
    torch.testing.assert_close(rv_b, rv_a, rtol=2e-6, atol=2e-6)
    # float64 reference of the statistics
    xd = x.double()
    mean, var = xd.mean(0), xd.var(0, unbiased=False)
    torch.testing.assert_close(cb[0].double(), mean, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(cb[1].double(), 1.0 / torch.sqrt(var + 1e-5), rtol=1e-5, atol=1e-6)
    unb = var * (M / max(M - 1, 1))
    torch.testing.assert_close(rv_b.double(), 0.9 * rv0.double() + 0.1 * unb, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("case", CASES[:6], ids=[f"m{m}_c{c}_p{p}" for m, c, p in CASES[:6]])
def test_bn_bwd_reduce_single_launch(case):
    from gan_danet_b200 import _lib as L, engine as E
    from gan_danet_b200._lib import ACT_RELU
    M, C, pitch = case
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(7 * M + C)
    x = torch.randn(M, pitch, generator=g).to(dev)[:, :C]
    dy = torch.randn(M, pitch, generator=g).to(dev)[:, :C]
    co = _coef(dev, C)
    co[0], co[1] = x.mean(0), 1.0 / torch.sqrt(x.var(0, unbiased=False) + 1e-5)
    co[2] = torch.randn(C, generator=g).to(dev) * co[1]
    co[3] = torch.randn(C, generator=g).to(dev) - co[0] * co[2]
    lib = E._lib(x)
    s_a = torch.empty(2 * C, dtype=torch.float64, device=dev)
    ws = E.workspace("stat", lib.gdn_stat_fused_ws_bytes(M, C), dev)
    L.check(lib.gdn_bn_bwd_reduce(dy.data_ptr(), pitch, 0, x.data_ptr(), pitch, 0, M, C, co[0].data_ptr(), co[1].data_ptr(), co[2].data_ptr(), co[3].data_ptr(),
                                  ACT_RELU, 0.0, s_a.data_ptr(), ws.data_ptr(), E._stream()), "gdn_bn_bwd_reduce")
    s_b = torch.empty(2 * C, dtype=torch.float64, device=dev)
    f_b = torch.empty(2 * C, dtype=torch.float32, device=dev)
    for _ in range(2):
        L.check(lib.gdn_bn_bwd_reduce_f(dy.data_ptr(), pitch, 0, x.data_ptr(), pitch, 0, M, C, co[0].data_ptr(), co[1].data_ptr(), co[2].data_ptr(), co[3].data_ptr(),
                                        ACT_RELU, 0.0, s_b.data_ptr(), f_b.data_ptr(), ws.data_ptr(), E.stat_counters(dev).data_ptr(), E._stream()), "gdn_bn_bwd_reduce_f")
    torch.cuda.synchronize()
    torch.testing.assert_close(s_b, s_a, rtol=1e-9, atol=1e-7)
    assert torch.equal(f_b, s_b.float())
    gate = ((x * co[2] + co[3]) > 0).double()
    gd = dy.double() * gate
    ref = torch.cat([gd.sum(0), (gd * (x.double() - co[0].double()) * co[1].double()).sum(0)])
    # the float64 gate differs from the kernel's fp32 fmaf gate on a few borderline elements of |dy| ~ 1 each
    torch.testing.assert_close(s_b, ref, rtol=1e-4, atol=1e-3 + 2e-5 * M)


def test_colsum_f32():
    from gan_danet_b200 import engine as E
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(32 * 128 * 256, 64, generator=g).to(dev)
    got = E.colsum_f32(x)
    torch.testing.assert_close(got.double(), x.double().sum(0), rtol=1e-5, atol=1e-3)
    r = torch.randn(5000, 1, generator=g).to(dev)
    torch.testing.assert_close(E.colsum_f32(r).double(), r.double().sum(0), rtol=1e-5, atol=1e-4)
