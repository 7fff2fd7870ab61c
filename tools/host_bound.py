"""Is the training step bound by the host (Python + ctypes launches) or by the GPU?  Prints the host enqueue time per step
(no synchronisation inside) next to the synchronised wall time per step."""
import os, sys, time
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
K = 5
t0 = time.perf_counter()
for _ in range(K):
    tr.train_step(*data)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3 * (t1 - t0) / K:.1f} ms/step   synchronised {1e3 * (t2 - t0) / K:.1f} ms/step")
