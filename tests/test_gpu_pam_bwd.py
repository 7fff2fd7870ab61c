"""Fused tcgen05 forward + backward of the position-attention core (pam_tc.cu) against a float64 autograd evaluation of
generator.py:115-122 on the same q, k, v -- at the REFERENCE's logit scale: no 1/sqrt(d) factor (generator.py:115-118), so the
logits of the trained/initialised network have std 8-11 and |max| 70-110.  q, k are drawn so that std(q.k) = 10.

Tolerances (relative L2, measured values in profiles/r02_pam_precision.json, tools/measure_pam_split.py):
  'fp16x3' (fp16 hi+lo split logit operands, the default): y <= 1e-3, dq / dk / dv <= 5e-3   (measured 6.4e-4, 2.8e-3, 2.4e-3, 1.4e-3 at N = 8192)
  'fp16'   (single fp16 logit operands):                   y <= 1.2e-3, dq / dk / dv <= 6e-3 (measured 8.2e-4, 3.7e-3, 3.5e-3, 2.0e-3)
Remaining error: bf16 softmax weights and values in the forward (the lazy-reference scheme needs bf16's exponent range, and kind::f16
takes one operand format per instruction), fp16 gradient operands in the backward.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(a, b):
    return float((a.detach().double() - b.detach().double()).norm() / b.detach().double().norm().clamp_min(1e-30))


def _reference(x, q, k, v, gamma, dy):
    """float64 evaluation one sample at a time (the N x N map of a sample is 512 MB in float64 at N = 8192)."""
    B, H, W, C = x.shape
    N, d = H * W, q.shape[-1]
    ys, dqs, dks, dvs, dg, stats = [], [], [], [], 0.0, []
    for b in range(B):
        qd, kd, vd = (t[b].double().reshape(N, -1).requires_grad_(True) for t in (q, k, v))
        gd = gamma.double().requires_grad_(True)
        S = qd @ kd.t()
        stats.append((float(S.detach().std()), float(S.detach().abs().max())))
        yref = gd * (torch.softmax(S, dim=-1) @ vd) + x[b].double().reshape(N, C)
        yref.backward(dy[b].double().reshape(N, C))
        ys.append(yref.detach()); dqs.append(qd.grad); dks.append(kd.grad); dvs.append(vd.grad); dg += float(gd.grad)
        del S, yref
    return torch.stack(ys), torch.stack(dqs), torch.stack(dks), torch.stack(dvs), dg, stats


@pytest.mark.parametrize("precision,tol_y,tol_g", [("fp16x3", 1e-3, 5e-3), ("fp16", 1.2e-3, 6e-3)])
@pytest.mark.parametrize("B,C,d,hw", [(2, 184, 23, (16, 32)), (1, 160, 20, (32, 32)), (2, 176, 22, (64, 128)), (2, 184, 23, (64, 128))])
def test_pam_core_reference_scale_logits(B, C, d, hw, precision, tol_y, tol_g):
    from gan_danet_b200 import engine as E
    H, W = hw
    N = H * W
    g = torch.Generator().manual_seed(B * 1000 + C)
    sig = (10.0 / d ** 0.5) ** 0.5                      # std(q.k) = sig^2 sqrt(d) = 10
    x = torch.randn(B, H, W, C, generator=g).to(DEV)
    q = (sig * torch.randn(B, H, W, d, generator=g)).to(DEV)
    k = (sig * torch.randn(B, H, W, d, generator=g)).to(DEV)
    v = torch.randn(B, H, W, C, generator=g).to(DEV)
    dy = (1e-4 * torch.randn(B, H, W, C, generator=g)).to(DEV)      # small, like a real mean-reduced loss gradient
    gamma = torch.full((1,), 0.5, device=DEV)
    assert E.pam_bwd_tensor_core
    tape = E.Tape()
    xv, qv, kv, vv, gv = E.Var(x), E.Var(q), E.Var(k), E.Var(v), E.Var(gamma)
    y = E.op_pam_core(tape, xv, qv, kv, vv, gv, precision=E.PAM_PRECISION_NAMES[precision])
    y.g = dy.clone()
    tape.backward()
    torch.cuda.synchronize()
    yref, dq, dk, dv, dg, stats = _reference(x, q, k, v, gamma, dy)
    assert all(8.0 < s[0] < 12.0 and s[1] > 50.0 for s in stats), stats          # the inputs really are at the reference's logit scale
    errs = {"y": rel(y.t.reshape(B, N, C), yref), "dq": rel(qv.g.reshape(B, N, d), dq), "dk": rel(kv.g.reshape(B, N, d), dk), "dv": rel(vv.g.reshape(B, N, C), dv)}
    assert errs["y"] < tol_y, errs
    assert max(errs["dq"], errs["dk"], errs["dv"]) < tol_g, errs
    assert abs(float(gv.g) - dg) < 2e-2 * abs(dg) + 1e-12, (float(gv.g), dg)      # dgamma = sum of N*C cancelling terms
    assert torch.equal(xv.g, dy)


def test_pam_core_gradient_scale_invariance():
    """The backward scales dy by a power of two into fp16's range (pam_tc.cu): gradients of 1e-9-sized and 1e+3-sized cotangents must be the same
    up to that factor -- bit for bit, because every scale involved is a power of two."""
    from gan_danet_b200 import engine as E
    B, H, W, C, d = 1, 16, 32, 184, 23
    g = torch.Generator().manual_seed(5)
    x, v, dy0 = (torch.randn(B, H, W, C, generator=g).to(DEV) for _ in range(3))
    q, k = ((1.4 * torch.randn(B, H, W, d, generator=g)).to(DEV) for _ in range(2))
    gamma = torch.full((1,), 0.5, device=DEV)

    def run(scale):
        tape = E.Tape()
        vs = [E.Var(t) for t in (x, q, k, v)]
        y = E.op_pam_core(tape, *vs, E.Var(gamma), precision=E.PREC_FP16X3)
        y.g = dy0 * scale
        tape.backward()
        torch.cuda.synchronize()
        return [t.g for t in vs[1:]]

    base = run(1.0)
    for p in (-30, 10):
        got = run(2.0 ** p)
        for a, b in zip(got, base):
            assert torch.equal(a, b * 2.0 ** p)
    zero = run(0.0)
    assert all(float(t.abs().max()) == 0.0 for t in zero)
