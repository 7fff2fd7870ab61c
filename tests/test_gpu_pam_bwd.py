"""Fused tcgen05 backward of the position-attention core (pam_tc.cu, two launches of pam_flash_bwd_kernel) against a
float64 autograd evaluation of generator.py:115-122 on the same q, k, v.

Operands are fp16 (logits) / bf16 (gradient-carrying GEMMs) with fp32 accumulation, so the tolerance is the bf16 operand
rounding (2^-9 per element, averaged over the contraction): 1e-2 relative L2 on dq, dk, dv, 2e-3 on y.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.mark.parametrize("B,C,d,hw,qs", [(2, 184, 23, (16, 32), 0.5), (1, 160, 20, (32, 32), 0.7), (2, 176, 22, (64, 128), 0.6)])
def test_pam_backward_tensor_core(B, C, d, hw, qs):
    from gan_danet_b200 import engine as E
    from gan_danet_b200._lib import PREC_FP16
    H, W = hw
    N = H * W
    g = torch.Generator().manual_seed(B * 1000 + C)
    x = torch.randn(B, H, W, C, generator=g).to(DEV)
    q = (qs * torch.randn(B, H, W, d, generator=g)).to(DEV)
    k = (qs * torch.randn(B, H, W, d, generator=g)).to(DEV)
    v = torch.randn(B, H, W, C, generator=g).to(DEV)
    dy = (1e-4 * torch.randn(B, H, W, C, generator=g)).to(DEV)      # small, like a real mean-reduced loss gradient
    gamma = torch.full((1,), 0.5, device=DEV)
    assert E.pam_bwd_tensor_core
    tape = E.Tape()
    xv, qv, kv, vv, gv = E.Var(x), E.Var(q), E.Var(k), E.Var(v), E.Var(gamma)
    y = E.op_pam_core(tape, xv, qv, kv, vv, gv, precision=PREC_FP16)
    y.g = dy.clone()
    tape.backward()
    torch.cuda.synchronize()

    qd, kd, vd = (t.double().reshape(B, N, -1).requires_grad_(True) for t in (q, k, v))
    gd = gamma.double().requires_grad_(True)
    P = torch.softmax(qd @ kd.transpose(1, 2), dim=-1)
    yref = gd * (P @ vd) + x.double().reshape(B, N, C)
    yref.backward(dy.double().reshape(B, N, C))
    assert rel(y.t.reshape(B, N, C), yref) < 2e-3
    assert rel(qv.g.reshape(B, N, d), qd.grad) < 1e-2, ("dq", rel(qv.g.reshape(B, N, d), qd.grad))
    assert rel(kv.g.reshape(B, N, d), kd.grad) < 1e-2, ("dk", rel(kv.g.reshape(B, N, d), kd.grad))
    assert rel(vv.g.reshape(B, N, C), vd.grad) < 1e-2, ("dv", rel(vv.g.reshape(B, N, C), vd.grad))
    assert abs(float(gv.g) - float(gd.grad)) < 1e-2 * abs(float(gd.grad)) + 1e-12
    assert torch.equal(xv.g, dy)
