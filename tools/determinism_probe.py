import os, sys
ROOT = "/root/repo"
sys.path.insert(0, ROOT)
import torch
import gan_danet_b200 as P
from gan_danet_b200 import engine as E
from gan_danet_b200.synthetic import make_batch
from gan_danet_b200.trainer import GANTrainer
dev = "cuda:0"
batches = [make_batch(10 * i, 2, 8, 16) for i in range(4)]
def run(conv, pam):
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
    perc = P.PerceptualLoss(pretrained=False, device=torch.device("cpu"))
    E.set_conv_precision(conv)
    G, D = G.to(dev), D.to(dev)
    perc.vgg.to(dev); perc.device = torch.device(dev)
    G.set_pam_precision(pam)
    tr = GANTrainer(G, D, perc, epochs=150); tr.epoch = 3
    res = []
    for i in range(6):
        out = tr.train_step(*(t.to(dev) for t in batches[i % 4]))
        res.append((float(out["loss_D"]), float(out["loss_G"])))
    return res
for conv, pam in (("fp32", "fp32"), ("bf16", "fp16")):
    a = run(conv, pam)
    junk = torch.randn(50_000_000, device=dev)   # perturb the allocator state between the runs
    b = run(conv, pam)
    print(conv, "identical" if a == b else "DIFFERENT", [f"{x[0]:.9g}/{y[0]:.9g}" for x, y in zip(a, b) if x != y][:3])
