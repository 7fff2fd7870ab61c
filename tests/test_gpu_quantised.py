"""The all-bf16 product mode (conv precision 'bf16' on tcgen05, fused PAM 'fp16x3', bf16-only operand shortcuts enabled: `bench.py --g-forward bf16`,
the benchmarked mode until the last session of round 2) asserted end to end against the quantisation-aware oracle (oracle/quantised_oracle.py: the reference algorithm, generator.py:230-247,
with rounding hooks exactly where these kernels round, accumulated in float64) -- SURVEY 7.4-1.

What the numbers mean:
* CUDA vs quantisation-aware oracle: the kernels compute the rounded algorithm they claim (same rounding points; fp32 accumulation in another
  summation order instead of float64).  SURVEY 7.4-1 hoped for <= 1e-3 here.  That is NOT attainable by any implementation, and the test shows why
  by running the oracle against ITSELF: the same rounded algorithm on the same CPU, accumulated in float32 instead of float64, lands 2.9e-3 (output)
  / 1.1e-1 (dx) away from its float64 run on this fixture.  Rounding to bf16 is a discontinuous map: two evaluations whose pre-rounding values differ
  by 1e-6 round a fraction of the activations to different neighbours (a 2^-9 jump each), 40 layers deep, and a ReLU whose pre-activation changes
  sign flips a whole unit's gradient -- a forward difference eps becomes a gradient difference ~sqrt(eps) in a ReLU/BN network
  (tools/precision_bisect.py on the CPU: rounding ONLY the gradient operands gives y 0 / dx 5e-3; rounding only the PAM core y 1.5e-4 / dx 2.6e-2;
  bf16x3 convolutions y 1.9e-5 / dx 3.9e-3).  The assertion is therefore: the CUDA path is as close to the oracle as the oracle is to itself
  (<= 2x the fp32-vs-fp64 self-distance), for output, dx and the whole parameter-gradient vector.  The per-kernel "same rounded operands"
  tests (tests/test_gpu_conv_tc.py, bitwise fused-edge tests in tests/test_gpu_parity.py) are where 1e-3 and tighter hold.
* quantisation-aware oracle vs the reference's float64 golden: the cost of the operand formats themselves; REPORTED (and bounded loosely), it is not
  an implementation property: y 9.0e-3, dx 1.9e-1 on this 8x16-grid / batch-2 fixture (BatchNorm over 256 values) for bf16 operands of ANY
  implementation, the reference's own modules included (SURVEY 7.4 measured 9.2e-3 / 1.8e-2..3.8e-2 on a 32x64 grid, batch 4).

The mode bench.py runs NOW (forward convolutions on hi+lo split operands: engine.generator_forward_x3) is the last test of this file: its output
is bitwise the parity mode's (1.5e-4 from the reference) and its gradients are within 2.7e-2 of the reference's float64 run.
"""
import json
import os
import sys

import pytest
import torch

from conftest import ROOT, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def qoracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import quantised_oracle
    return quantised_oracle


def _seeded_generator(seed, gamma):
    import gan_danet_b200 as P
    from gan_danet_b200.models.generator import CAMModule, PAMModule
    torch.manual_seed(seed)
    G = P.FlexibleUpsamplingModule(46)
    G.apply(P.weights_init_normal)
    with torch.no_grad():
        for m in G.modules():
            if isinstance(m, (PAMModule, CAMModule)):
                m.gamma.fill_(gamma)
    return G.train()


def _oracle_run(qoracle, G, x, r, fmt, dtype=torch.float64):
    pnames = [k for k, _ in G.named_parameters()]
    sd = {k: (v.detach().to(dtype).cpu() if v.is_floating_point() else v.detach().cpu()) for k, v in G.state_dict().items()}
    sdp = {k: (v.clone().requires_grad_(True) if k in pnames else v) for k, v in sd.items()}
    xd = x.detach().clone().to(dtype).requires_grad_(True)        # clone: .to() of a tensor already in `dtype` would alias the golden fixture
    bufs = {}
    y = qoracle.generator_forward(sdp, xd, fmt, training=True, buffers_out=bufs)
    grads = torch.autograd.grad((y * r.to(dtype)).sum(), [xd] + [sdp[k] for k in pnames], allow_unused=True)
    return y.detach(), grads[0], {k: (g if g is not None else torch.zeros_like(sdp[k])) for k, g in zip(pnames, grads[1:])}, bufs


def _whole_vector(grads, ref, skip=("key.bias",)):
    keys = [k for k in ref if not any(s in k for s in skip)]
    num = sum(float((grads[k].detach().double().cpu() - ref[k].double()).norm() ** 2) for k in keys)
    den = sum(float(ref[k].double().norm() ** 2) for k in keys)
    return (num / den) ** 0.5


@pytest.mark.parametrize("pam", ["fp16x3", "fp16"])
def test_generator_product_mode_vs_quantised_oracle(golden, qoracle, pam):
    from gan_danet_b200 import engine as E
    g = golden("generator_cin46_8x16")
    G = _seeded_generator(g["seed"], g["gamma"])
    fmt = qoracle.Formats(pam_logits=None if pam == "fp16x3" else "fp16")
    yq, dxq, gq, bufq = _oracle_run(qoracle, G, g["x"], g["r"], fmt)
    y32, dx32, g32, _ = _oracle_run(qoracle, G, g["x"], g["r"], fmt, dtype=torch.float32)      # the oracle against itself: float32 accumulation
    self_d = {"y": rel_err(y32, yq), "dx": rel_err(dx32, dxq), "grads_whole_vector": _whole_vector(g32, gq)}
    old = E.conv_precision
    E.set_conv_precision("bf16")
    try:
        assert E.bf16_storage_ok() and E.fuse_bn_into_pack and E.conv_bn_packed_grad and E.danet_cat16 and E.pam_v16_from_conv      # the bench configuration
        G.set_pam_precision(pam)
        Gd = G.to(DEV)
        x = g["x"].to(DEV).requires_grad_(True)
        y = Gd(x)
        y.backward(g["r"].to(DEV))
        torch.cuda.synchronize()
    finally:
        E.set_conv_precision(old)
    grads = {k: p.grad for k, p in Gd.named_parameters()}
    rep = {"quantised_oracle_fp32_vs_fp64_self_distance": self_d,
           "cuda_vs_quantised_oracle": {"y": rel_err(y, yq), "dx": rel_err(x.grad, dxq), "grads_whole_vector": _whole_vector(grads, gq)},
           "quantised_oracle_vs_fp64_reference": {"y": rel_err(yq, g["y"]), "dx": rel_err(dxq, g["dx"]),
                                                  "grads_whole_vector": _whole_vector(gq, {k: v for k, v in g["grads_small"].items()})},
           "cuda_vs_fp64_reference": {"y": rel_err(y, g["y"]), "dx": rel_err(x.grad, g["dx"])}}
    print("\n[precision] bf16 convs + PAM %s: %s" % (pam, json.dumps(rep)))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rep, open(os.path.join(ROOT, "gpurun_out", f"precision_product_mode_{pam}.json"), "w"), indent=1)
    c = rep["cuda_vs_quantised_oracle"]
    for k in ("y", "dx", "grads_whole_vector"):                          # as close to the rounded algorithm as that algorithm is to itself
        assert c[k] < max(1e-3, 2.0 * self_d[k]), (k, rep)
    q = rep["quantised_oracle_vs_fp64_reference"]
    assert q["y"] < 2e-2 and rep["cuda_vs_fp64_reference"]["y"] < 2e-2, rep     # cost of bf16 operands (9.2e-3), reported
    # BN running statistics after the step: a mean / variance over the rounded activations
    sd = Gd.state_dict()
    for k, v in bufq.items():
        if "running" in k:
            assert rel_err(sd[k], v) < 5e-3, (k, rel_err(sd[k], v))


def test_generator_parity_mode_with_fused_pam(golden):
    """'bf16x3' convolutions (hi+lo split, ~fp32) + fused PAM 'fp16x3': the tensor-core mode for parity against the reference's float64 run.
    Output <= 1e-3 (north_star).  Gradients: bf16x3 convs with the fp32 PAM measure dx 3.9e-3 on this fixture (test_generator), the sqrt law of a
    1.9e-5 output difference; the fused PAM adds its own 1.5e-4 output difference (bf16 softmax weights / values) => asserted at 4e-2, the
    round-1 fp16 kernel's 5e-2 row is kept in test_generator for comparison."""
    from gan_danet_b200 import engine as E
    g = golden("generator_cin46_8x16")
    G = _seeded_generator(g["seed"], g["gamma"])
    G.set_pam_precision("fp16x3")
    old = E.conv_precision
    E.set_conv_precision("bf16x3")
    try:
        Gd = G.to(DEV)
        x = g["x"].to(DEV).requires_grad_(True)
        y = Gd(x)
        y.backward(g["r"].to(DEV))
        torch.cuda.synchronize()
    finally:
        E.set_conv_precision(old)
    grads = {k: p.grad for k, p in Gd.named_parameters()}
    ey, edx, eg = rel_err(y, g["y"]), rel_err(x.grad, g["dx"]), _whole_vector(grads, g["grads_small"])
    print(f"\n[precision] bf16x3 convs + PAM fp16x3 vs fp64 reference: y {ey:.2e} dx {edx:.2e} grads {eg:.2e}")
    assert ey < 1e-3, ey
    assert edx < 4e-2 and eg < 4e-2, (edx, eg)


def test_generator_benchmarked_mode_forward_x3(golden, qoracle):
    """The mode bench.py runs: conv precision 'bf16' with engine.generator_forward_x3 -- the generator's FORWARD convolutions on hi+lo split operands
    (recorded under conv_precision_scope('bf16x3'), like Discriminator1's), every gradient GEMM on single bf16 operands (the hi parts) -- fused PAM
    'fp16x3'.
    * the forward is the parity mode's: bitwise the output of conv precision 'bf16x3' (same kernels, same operands) and <= 1e-3 from the reference's
      float64 run (north_star's bar on the generator output);
    * the gradients are those of a network whose forward is exact to 1.5e-4 and whose backward operands are bf16: the quantisation-aware oracle with
      Formats.forward_x3() predicts dx 4.1e-2 / parameter gradients 3.8e-2 against the float64 reference (bf16 forward: 1.9e-1 / 1.8e-1 -- flipped ReLU
      masks); asserted <= 1e-1 against the reference and against the oracle."""
    from gan_danet_b200 import engine as E
    g = golden("generator_cin46_8x16")

    def run(conv, gx3, shortcuts=True):
        G = _seeded_generator(g["seed"], g["gamma"])
        G.set_pam_precision("fp16x3")
        old, old_gx3, old_sc = E.conv_precision, E.generator_forward_x3, (E.pam_v16_from_conv, E.conv_bn_packed_grad)
        E.set_conv_precision(conv)
        E.generator_forward_x3 = gx3
        E.pam_v16_from_conv = E.conv_bn_packed_grad = shortcuts
        try:
            assert E.generator_forward_precision() == ("bf16x3" if gx3 else None)
            Gd = G.to(DEV)
            x = g["x"].to(DEV).requires_grad_(True)
            y = Gd(x)
            assert E.conv_precision == conv                       # the scope restored the mode
            y.backward(g["r"].to(DEV))
            torch.cuda.synchronize()
        finally:
            E.set_conv_precision(old)
            E.generator_forward_x3 = old_gx3
            E.pam_v16_from_conv, E.conv_bn_packed_grad = old_sc
        return G, y.detach(), x.grad, {k: p.grad for k, p in Gd.named_parameters()}, Gd.state_dict()

    _, y_par, _, _, _ = run("bf16x3", False)
    G, y, dx, grads, sd = run("bf16", True)
    assert torch.equal(y, y_par), rel_err(y, y_par)
    # the two operand shortcuts that stay on inside the split forward (engine.split_forward_in_product_mode: the value projection's epilogue emits the
    # fused PAM's bf16 V operand; BatchNorm's backward hands the convolution its dz as a bf16 operand only) change no bit of the output or of any gradient
    _, y_ns, dx_ns, grads_ns, _ = run("bf16", True, shortcuts=False)
    assert torch.equal(y, y_ns) and torch.equal(dx, dx_ns)
    assert not [k for k in grads if not torch.equal(grads[k], grads_ns[k])]
    yq, dxq, gq, bufq = _oracle_run(qoracle, G, g["x"], g["r"], qoracle.Formats.forward_x3())
    rep = {"cuda_vs_fp64_reference": {"y": rel_err(y, g["y"]), "dx": rel_err(dx, g["dx"]), "grads_whole_vector": _whole_vector(grads, g["grads_small"])},
           "cuda_vs_quantised_oracle": {"y": rel_err(y, yq), "dx": rel_err(dx, dxq), "grads_whole_vector": _whole_vector(grads, gq)},
           "quantised_oracle_vs_fp64_reference": {"y": rel_err(yq, g["y"]), "dx": rel_err(dxq, g["dx"]),
                                                  "grads_whole_vector": _whole_vector(gq, {k: v for k, v in g["grads_small"].items()})}}
    print("\n[precision] bf16 convs, generator forward x3 + PAM fp16x3: %s" % json.dumps(rep))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rep, open(os.path.join(ROOT, "gpurun_out", "precision_benchmarked_mode_gx3.json"), "w"), indent=1)
    c = rep["cuda_vs_fp64_reference"]
    assert c["y"] < 1e-3, rep
    assert c["dx"] < 1e-1 and c["grads_whole_vector"] < 1e-1, rep
    q = rep["cuda_vs_quantised_oracle"]
    assert q["y"] < 1e-3 and q["dx"] < 1e-1 and q["grads_whole_vector"] < 1e-1, rep
    for k, v in g["buffers_after"].items():
        assert rel_err(sd[k], v) < 1e-4, (k, rel_err(sd[k], v))
