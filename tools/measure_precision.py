"""Prints the generator's output / gradient errors against the float64 golden fixture for every conv / PAM precision mode."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import torch
from conftest import rel_err, ROOT
import gan_danet_b200 as P
from gan_danet_b200 import engine as E
from gan_danet_b200.models.generator import CAMModule, PAMModule

g = torch.load(os.path.join(ROOT, "tests", "golden", "generator_cin46_8x16.pt"), weights_only=False)
for conv in ("fp32", "bf16x3", "bf16"):
    for pam in ("fp32", "fp16", "fp16x3"):
        E.set_conv_precision(conv)
        torch.manual_seed(g["seed"])
        G = P.FlexibleUpsamplingModule(46); G.apply(P.weights_init_normal)
        with torch.no_grad():
            for m in G.modules():
                if isinstance(m, (PAMModule, CAMModule)): m.gamma.fill_(g["gamma"])
        G.set_pam_precision(pam); G = G.train().to("cuda:0")
        x = g["x"].to("cuda:0").requires_grad_(True)
        y = G(x); y.backward(g["r"].to("cuda:0")); torch.cuda.synchronize()
        grads = {k: p.grad for k, p in G.named_parameters()}
        num = sum(float((grads[k].double().cpu() - v.double()).norm() ** 2) for k, v in g["grads_small"].items() if "key.bias" not in k)
        den = sum(float(v.double().norm() ** 2) for k, v in g["grads_small"].items() if "key.bias" not in k)
        print(f"conv={conv:7s} pam={pam}: y {rel_err(y, g['y']):.2e}  dx {rel_err(x.grad, g['dx']):.2e}  small-grads(whole vector) {(num/den)**0.5:.2e}", flush=True)
