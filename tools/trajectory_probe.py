"""Free-running loss trajectory of the CUDA engine (fp32 / bf16x3 / bf16 convs) against the float64 CPU oracle from identical weights:
prints the relative deviation of loss_D / loss_G per step (tiny shapes: B = 2, grid 8x16 -> 32x64)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import gan_danet_oracle as oracle
import gan_danet_b200 as P
from gan_danet_b200 import engine as E
from gan_danet_b200.synthetic import make_batch
from gan_danet_b200.trainer import GANTrainer

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
dev = "cuda:0"
batches = [make_batch(10 * i, 2, 8, 16) for i in range(4)]


def fresh():
    torch.manual_seed(11)
    G, D = P.FlexibleUpsamplingModule(46), P.Discriminator1()
    G.apply(P.weights_init_normal)
    for mod in (D.conv1, D.conv2, D.conv3, D.conv4, D.fc2):
        mod.apply(P.weights_init_normal)
    D._materialise_fc1(batches[0][1])
    with torch.no_grad():
        for n, p in G.named_parameters():
            if n.endswith("gamma"):
                p.fill_(0.05)
    torch.manual_seed(12)
    return G, D, P.PerceptualLoss(pretrained=False, device=torch.device("cpu"))


G, D, perc = fresh()
d64 = lambda sd: {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
st = oracle.TrainState(d64(G.state_dict()), d64(D.state_dict()), {k: v.double() for k, v in perc.vgg.state_dict().items()})
ref = [oracle.train_step(st, *(t.double() for t in batches[i % 4]), epoch=3, epochs=150) for i in range(steps)]
for conv, pam in (("fp32", "fp32"), ("bf16x3", "fp16"), ("bf16", "fp16")):
    G, D, perc = fresh()
    E.set_conv_precision(conv)
    G, D = G.to(dev), D.to(dev)
    perc.vgg.to(dev); perc.device = torch.device(dev)
    G.set_pam_precision(pam)
    tr = GANTrainer(G, D, perc, epochs=150); tr.epoch = 3
    worst = []
    for i in range(steps):
        out = tr.train_step(*(t.to(dev) for t in batches[i % 4]))
        worst.append(max(abs(float(out[k]) - ref[i][k]) / max(abs(ref[i][k]), 1e-3) for k in ("loss_D", "loss_G")))
    marks = [0, 1, 2, 4, 7, 9, 14, 19, 29]
    print(f"conv {conv:6s} pam {pam}: " + "  ".join(f"step {m}: {worst[m]:.1e}" for m in marks if m < steps), flush=True)
