"""One G+D training step at the bench shapes inside a cudaProfilerStart/Stop window (for `ncu --profile-from-start off`).

    python tools/profile_step.py [--batch 32] [--grid 64x128] [--pam-only]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import gan_danet_b200 as P  # noqa: E402
from gan_danet_b200.synthetic import fast_batch  # noqa: E402
from gan_danet_b200.trainer import GANTrainer, init_like_reference  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--grid", default="64x128")
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--conv-precision", default="bf16")
ap.add_argument("--g-forward", default="x3", choices=["x3", "bf16"], help="as bench.py: the generator's forward convolutions on hi+lo split operands (default) or single bf16")
ap.add_argument("--pam-only", action="store_true", help="profile only a PAM forward (C=184) instead of the whole step")
args = ap.parse_args()
h, w = (int(v) for v in args.grid.split("x"))
dev = torch.device("cuda:0")
from gan_danet_b200 import engine as E  # noqa: E402
E.set_conv_precision(args.conv_precision)
E.generator_forward_x3 = args.g_forward == "x3" and args.conv_precision == "bf16"
torch.manual_seed(0)
if args.pam_only:
    from gan_danet_b200.models.generator import PAMModule
    m = PAMModule(184)
    m.apply(P.weights_init_normal)
    with torch.no_grad():
        m.gamma.fill_(0.5)
    m = m.to(dev)
    x = 0.5 * torch.randn(args.batch, 184, h, w, device=dev)
    with torch.no_grad():
        for _ in range(args.warmup):
            m(x)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        m(x)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
else:
    G, D = P.FlexibleUpsamplingModule(46), P.Discriminator1()
    lr05, real, aux = fast_batch(1000, args.batch, h, w)
    init_like_reference(G, D, real)
    with torch.no_grad():
        for n, p in G.named_parameters():
            if n.endswith("gamma"):
                p.fill_(0.05)
    torch.manual_seed(2)
    perc = P.PerceptualLoss(pretrained=False, device=torch.device("cpu"))
    perc.vgg.to(dev)
    perc.device = dev
    G, D = G.to(dev), D.to(dev)
    tr = GANTrainer(G, D, perc)
    tr.epoch = 3
    data = [t.to(dev) for t in (lr05, real, aux)]
    for _ in range(args.warmup):
        tr.train_step(*data)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    out = tr.train_step(*data)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("loss_D", float(out["loss_D"]), "loss_G", float(out["loss_G"]))
print("done")
