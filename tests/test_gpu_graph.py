"""The training step captured in one CUDA graph (trainer.GraphedTrainStep) against the eager step: same kernels, same order, same scalars =>
bit-identical losses, generated fields, parameters and AdamW state, step after step, including across an epoch change (adversarial weight and
learning rates live in device memory, not in the captured kernel arguments)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _setup(conv):
    import gan_danet_b200 as P
    from gan_danet_b200.synthetic import make_batch
    from gan_danet_b200.trainer import GANTrainer
    batches = [tuple(t.to(DEV) for t in make_batch(10 * i, 2, 8, 16)) for i in range(3)]
    torch.manual_seed(11)
    G = P.FlexibleUpsamplingModule(46)
    D = P.Discriminator1()
    G.apply(P.weights_init_normal)
    D.apply(P.weights_init_normal)
    D._materialise_fc1(batches[0][1].cpu())
    with torch.no_grad():
        for n, p in G.named_parameters():
            if n.endswith("gamma"):
                p.fill_(0.05)
    torch.manual_seed(12)
    perc = P.PerceptualLoss(pretrained=False, device=torch.device("cpu"))

    def make():
        g, d, pl = copy.deepcopy(G).to(DEV), copy.deepcopy(D).to(DEV), P.PerceptualLoss(pretrained=False, device=torch.device("cpu"))
        pl.vgg.load_state_dict(perc.vgg.state_dict())
        pl.vgg.to(DEV)
        pl.device = torch.device(DEV)
        tr = GANTrainer(g, d, pl, epochs=150)
        tr.epoch = 3
        return tr

    return batches, make


@pytest.mark.parametrize("conv", ["bf16", "fp32"])
def test_graphed_step_is_bit_identical_to_eager(conv):
    from gan_danet_b200 import engine as E
    from gan_danet_b200.trainer import GraphedTrainStep
    old = E.conv_precision
    E.set_conv_precision(conv)
    try:
        batches, make = _setup(conv)
        eager, graphed = make(), make()
        warm = 2
        for _ in range(warm):
            eager.train_step(*batches[0])
        step = GraphedTrainStep(graphed, *batches[0], warmup=warm)
        assert step.launches_per_step > 100
        for i in range(6):
            if i == 3:                     # epoch boundary: w = epoch / epochs and both learning rates change (GAN_DANet_train.ipynb:266,294-295)
                eager.end_epoch()
                graphed.end_epoch()
            b = batches[i % 3]
            ref = eager.train_step(*b)
            out = step(*b)
            torch.cuda.synchronize()
            for k in ("loss_D", "loss_G", "adv", "pixel", "tv", "perceptual", "hr"):
                assert torch.equal(out[k], ref[k]), (i, k, float((out[k] - ref[k]).abs().max()))
        for (n, p), (_, q) in zip(eager.G.named_parameters(), graphed.G.named_parameters()):
            assert torch.equal(p, q), n
        for (n, p), (_, q) in zip(eager.D.named_parameters(), graphed.D.named_parameters()):
            assert torch.equal(p, q), n
        for (k, v), (_, u) in zip(eager.G.state_dict().items(), graphed.G.state_dict().items()):
            assert torch.equal(v, u), k            # BatchNorm running statistics and num_batches_tracked
        pe, pg = next(iter(eager.D.parameters())), next(iter(graphed.D.parameters()))
        assert torch.equal(eager.opt_D.state[pe]["exp_avg_sq"], graphed.opt_D.state[pg]["exp_avg_sq"])
        assert eager.opt_D.state[pe]["step"] == graphed.opt_D.state[pg]["step"] == warm + 6
    finally:
        E.set_conv_precision(old)
        E.release_buffers()
