"""TF32 tcgen05 linear-layer kernels (linear_tc.cu: Discriminator1.fc1, discriminator.py:66,75) against float64.
TF32 operands (10-bit mantissa) with fp32 accumulation over K >= 8192 terms: 2e-3 relative L2."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


# (16, 128, .): the weight gradient stays on the fp32 engine (its tensor-core form needs Mb % 32 == 0 and N % 256 == 0)
@pytest.mark.parametrize("Mb,N,K", [(16, 128, 8192), (64, 256, 16384), (32, 1024, 32768)])
def test_linear_tc(Mb, N, K):
    from gan_danet_b200 import engine as E
    from gan_danet_b200._lib import ACT_LRELU
    g = torch.Generator().manual_seed(Mb + N)
    x = torch.randn(Mb, K, generator=g).to(DEV)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    r = torch.randn(Mb, N, generator=g).to(DEV)
    old = E.linear_tensor_core
    E.linear_tensor_core = True
    try:
        tape = E.Tape()
        xv, wv, bv = E.Var(x), E.Var(w), E.Var(b)
        y = E.op_linear(tape, xv, wv, bv)          # no activation here: a sign flip of y ~ 0 under TF32 would change dz itself
        y.g = r.clone()
        tape.backward()
        ya = E.op_linear(E.Tape(record=False), E.Var(x), E.Var(w), E.Var(b), act=ACT_LRELU, slope=0.2)
        torch.cuda.synchronize()
    finally:
        E.linear_tensor_core = old
    xd, wd, bd = x.double().requires_grad_(True), w.double().requires_grad_(True), b.double().requires_grad_(True)
    yref = xd @ wd.t() + bd
    yref.backward(r.double())
    assert rel(ya.t, torch.nn.functional.leaky_relu(yref.detach(), 0.2)) < 2e-3
    errs = {"y": rel(y.t, yref), "dx": rel(xv.g, xd.grad), "dw": rel(wv.g, wd.grad)}
    assert all(e < 2e-3 for e in errs.values()), errs
    assert rel(bv.g, bd.grad) < 1e-5
