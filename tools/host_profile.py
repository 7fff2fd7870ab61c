"""cProfile of the host side of the training step (what the 48 ms of enqueue time per step consist of)."""
import cProfile, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gan_danet_b200 as P
from gan_danet_b200 import engine as E
from gan_danet_b200.synthetic import fast_batch
from gan_danet_b200.trainer import GANTrainer, init_like_reference
dev = torch.device("cuda:0")
E.set_conv_precision("bf16")
torch.manual_seed(0)
G, D = P.FlexibleUpsamplingModule(46), P.Discriminator1()
lr05, real, aux = fast_batch(1000, 32, 64, 128)
init_like_reference(G, D, real)
torch.manual_seed(2)
perc = P.PerceptualLoss(pretrained=False, device=torch.device("cpu"))
perc.vgg.to(dev); perc.device = dev
G, D = G.to(dev), D.to(dev)
tr = GANTrainer(G, D, perc); tr.epoch = 3
data = [t.to(dev) for t in (lr05, real, aux)]
for _ in range(3):
    tr.train_step(*data)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    tr.train_step(*data)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
