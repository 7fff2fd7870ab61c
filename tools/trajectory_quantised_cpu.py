"""CPU only: what bf16 operands cost the LOSSES of the teacher-forced walk, independently of any CUDA code.

tests/test_gpu_trajectory.py walks 200 teacher-forced steps and finds the benchmarked mode (bf16 convolution operands) above 1 % on 3 of 400 loss
values (worst 1.2 %).  Is that this implementation, or the format?  This tool walks the same steps with the CPU oracle (same seeds, schedule, batches)
and at every state evaluates the step's forward losses a second time with the QUANTISATION-AWARE oracle's generator (oracle/quantised_oracle.py: the
reference algorithm, float64 accumulation, operands rounded exactly where the product-mode kernels round) -- everything else (Discriminator1, VGG19,
the losses) in float64 and UNROUNDED, i.e. a lower bound of what a bf16-operand step can achieve:

    loss_D'  = (BCE(D(real), 1) + BCE(D(G_q(x)), 0)) / 2                       D before its update, as in the step
    loss_G'  = (1 - w) MSE + w BCE(D_updated(G_q(x)), 1) + TV + perceptual      D_updated = the reference walk's updated discriminator

and reports |loss' - loss| / |loss| per step.  Test infrastructure (imports oracle/).

    python tools/trajectory_quantised_cpu.py [--steps 200] [--json out.json]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gan_danet_oracle as O  # noqa: E402
import quantised_oracle as Q  # noqa: E402
import gan_danet_b200 as P  # noqa: E402  (module mirror: only used on the CPU to build the seeded initial state, as the GPU test does)
from gan_danet_b200.synthetic import make_batch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--json", default=None)
args = ap.parse_args()

epochs, per_epoch = 150, 4
batches = [make_batch(10 * i, 2, 8, 16) for i in range(per_epoch)]
torch.manual_seed(11)
G0, D0 = P.FlexibleUpsamplingModule(46), P.Discriminator1()
G0.apply(P.weights_init_normal)
for mod in (D0.conv1, D0.conv2, D0.conv3, D0.conv4, D0.fc2):
    mod.apply(P.weights_init_normal)
D0._materialise_fc1(batches[0][1])
with torch.no_grad():
    for n, p in G0.named_parameters():
        if n.endswith("gamma"):
            p.fill_(0.05)
torch.manual_seed(12)
vgg_sd = {k: v.clone() for k, v in P.PerceptualLoss(pretrained=False, device=torch.device("cpu")).vgg.state_dict().items()}
st = O.TrainState({k: v.clone() for k, v in G0.state_dict().items()}, {k: v.clone() for k, v in D0.state_dict().items()}, vgg_sd)
vgg64 = {k: v.double() for k, v in vgg_sd.items()}
d64 = lambda sd: {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}  # noqa: E731

fmts = {"bf16 operands (product mode)": Q.Formats(),
        # the benchmarked mode (engine.generator_forward_x3): forward convolutions on hi+lo split operands; what remains is the fused PAM's bf16 P / V
        "split forward operands (benchmarked mode: conv x / w hi+lo, PAM P / V bf16)": Q.Formats.forward_x3(),
        # what-if (not implemented on the device): tcgen05 kind::f16 takes fp16 operands at the bf16 rate; activations and weights of this network fit fp16's range
        "fp16 forward operands (what-if: conv x / w and PAM P / V in fp16, gradient operands unchanged)": Q.Formats("fp16", "fp16", "bf16", "fp16", "fp16", "fp16", None),
        "exact (float64 re-evaluation: the float32 oracle's own noise)": Q.Formats.exact()}
log = {k: [] for k in fmts}
t0 = time.time()
with torch.no_grad():
    pass
for i in range(args.steps):
    epoch = i // per_epoch
    lr_d, lr_g = O.cosine_warm_restarts_lr(epoch, 4e-4), O.cosine_warm_restarts_lr(epoch, 2e-4)
    lr05, real, aux = batches[i % per_epoch]
    g_before, d_before = d64(st.g), d64(st.d)
    ref = O.train_step(st, lr05, real, aux, epoch=epoch, epochs=epochs, lr_d=lr_d, lr_g=lr_g)       # st: post-step state
    d_after = d64(st.d)
    w = epoch / epochs
    with torch.no_grad():
        x = O.prepare_input(lr05.double(), aux.double())
        r64 = real.double()
        for name, f in fmts.items():
            hr = Q.generator_forward(g_before, x, f, training=True)
            loss_d = (O.bce_with_logits(O.discriminator_forward(d_before, r64), 1.0) + O.bce_with_logits(O.discriminator_forward(d_before, hr), 0.0)) / 2
            loss_g = (1 - w) * O.mse(hr, r64) + w * O.bce_with_logits(O.discriminator_forward(d_after, hr), 1.0) + O.tv_loss(hr, 1e-5) + O.perceptual_loss(vgg64, hr, r64)
            log[name].append({"step": i, "loss_D": abs(float(loss_d) - ref["loss_D"]) / abs(ref["loss_D"]), "loss_G": abs(float(loss_g) - ref["loss_G"]) / abs(ref["loss_G"])})
    if i % 20 == 19:
        print(f"step {i + 1}: " + "; ".join(f"{k.split(' ')[0]} worst so far D {max(r['loss_D'] for r in v):.2e} G {max(r['loss_G'] for r in v):.2e}" for k, v in log.items()),
              f"({time.time() - t0:.0f} s)", flush=True)

out = {}
for name, v in log.items():
    d, g = sorted(r["loss_D"] for r in v), sorted(r["loss_G"] for r in v)
    out[name] = {"steps": len(v), "loss_D": {"median": d[len(d) // 2], "worst": d[-1], "above_1pct": sum(x > 1e-2 for x in d)},
                 "loss_G": {"median": g[len(g) // 2], "worst": g[-1], "above_1pct": sum(x > 1e-2 for x in g)}}
    print(name, json.dumps(out[name]))
if args.json:
    json.dump({"what": "tools/trajectory_quantised_cpu.py: relative deviation of the teacher-forced step losses when ONLY the generator forward uses the product mode's "
                       "operand formats (quantisation-aware oracle, float64 accumulate; D, VGG19 and the losses unrounded float64), against the float32 CPU oracle walk",
               "summary": out, "per_step": log}, open(args.json, "w"), indent=1)
