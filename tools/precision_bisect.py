"""CPU-only: distance of the quantisation-aware oracle (oracle/quantised_oracle.py) to the reference's float64 run of the generator
(tests/golden/generator_cin46_8x16.pt), one rounding point at a time.  It answers which operand format costs what in the tensor-core
product mode without a GPU.  Test infrastructure (imports oracle/).

    python tools/precision_bisect.py [--layers] [--json out.json]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gan_danet_oracle as O  # noqa: E402
import quantised_oracle as Q  # noqa: E402
import gan_danet_b200 as P  # noqa: E402  (module mirror: only used on the CPU to build the seeded state_dict)


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def seeded_state(seed, gamma):
    from gan_danet_b200.models.generator import CAMModule, PAMModule
    torch.manual_seed(seed)
    G = P.FlexibleUpsamplingModule(46)
    G.apply(P.weights_init_normal)
    with torch.no_grad():
        for m in G.modules():
            if isinstance(m, (PAMModule, CAMModule)):
                m.gamma.fill_(gamma)
    return {k: v.double() for k, v in G.state_dict().items()}, [k for k, _ in G.named_parameters()]


def run(g, f, fmt_for=None):
    sd, pnames = seeded_state(g["seed"], g["gamma"])
    sdp = {k: (v.clone().requires_grad_(True) if k in pnames else v) for k, v in sd.items()}
    x = g["x"].double().requires_grad_(True)
    y = Q.generator_forward(sdp, x, f, fmt_for=fmt_for)
    grads = torch.autograd.grad((y * g["r"].double()).sum(), [x] + [sdp[k] for k in pnames], allow_unused=True)
    gd = {k: (t if t is not None else torch.zeros_like(sdp[k])) for k, t in zip(pnames, grads[1:])}
    return y.detach(), grads[0], gd


def summarise(g, y, dx, gd):
    small = {k: v for k, v in g["grads_small"].items() if "key.bias" not in k}
    num = sum(float((gd[k] - v.double()).norm() ** 2) for k, v in small.items())
    den = sum(float(v.double().norm() ** 2) for v in small.values())
    return {"y": rel(y, g["y"]), "dx": rel(dx, g["dx"]), "grads_whole_vector": (num / den) ** 0.5}


if __name__ == "__main__":
    g = torch.load(os.path.join(ROOT, "tests", "golden", "generator_cin46_8x16.pt"), weights_only=False)
    cases = {
        "exact (must reproduce the golden)": Q.Formats.exact(),
        "all-bf16 product mode: bf16 convs + fp16x3 PAM": Q.Formats(),
        "BENCHMARKED mode: forward convolutions on hi+lo split operands, gradient GEMMs on bf16 (engine.generator_forward_x3)": Q.Formats.forward_x3(),
        "product mode with single-fp16 logits (PAM 'fp16')": Q.Formats(pam_logits="fp16"),
        "what-if, not implemented: fp16 forward operands (conv x / w, PAM P / V), bf16 conv gradient operand": Q.Formats("fp16", "fp16", "bf16", "fp16", "fp16", "fp16", None),
        "conv operands only (x, w bf16; gradient operand exact)": Q.Formats("bf16", "bf16", None, None, None, None, None),
        "conv gradient operand only": Q.Formats(None, None, "bf16", None, None, None, None),
        "conv x only": Q.Formats("bf16", None, None, None, None, None, None),
        "conv w only": Q.Formats(None, "bf16", None, None, None, None, None),
        "PAM core only (bf16 P, V; fp16 backward)": Q.Formats(None, None, None, "bf16", "bf16", "fp16", None),
        "PAM core only, round-1 formats (fp16 logits, bf16 backward)": Q.Formats(None, None, None, "bf16", "bf16", "bf16", "fp16"),
    }
    out = {}
    for name, f in cases.items():
        res = summarise(g, *run(g, f))
        out[name] = res
        print(f"{name:118s} y {res['y']:.2e}  dx {res['dx']:.2e}  grads {res['grads_whole_vector']:.2e}", flush=True)
    if "--layers" in sys.argv:
        # which layer group's bf16 operand rounding costs what: the product mode's formats on ONE group of convolutions, exact elsewhere
        groups = {"initial conv": lambda n: n == "initial", "dense-layer convs": lambda n: n.startswith("dense"), "PAM (projections + core)": lambda n: n.startswith("pam"),
                  "fuse convs": lambda n: n.startswith("fuse"), "transition convs": lambda n: n.startswith("trans"), "upsample conv 0": lambda n: n == "up0",
                  "upsample conv 1": lambda n: n == "up1", "skip projections": lambda n: n.startswith("adjust")}
        ex, pr = Q.Formats.exact(), Q.Formats()
        for name, pred in groups.items():
            res = summarise(g, *run(g, pr, fmt_for=lambda n, pred=pred: pr if pred(n) else ex))
            out["only " + name] = res
            print(f"{'only ' + name:70s} y {res['y']:.2e}  dx {res['dx']:.2e}  grads {res['grads_whole_vector']:.2e}", flush=True)
    if "--json" in sys.argv:
        json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
