"""Drop-in test: the notebook's own loop (GAN_DANet_train.ipynb:182-194,225-269, restated in oracle/notebook_step.py with STOCK torch.optim.AdamW,
nn.BCEWithLogitsLoss(), nn.MSELoss(), F.interpolate for the input preparation and D(hr_generated) with D's parameters requiring grad in the
generator step) runs UNCHANGED on this repo's modules -- `import gan_danet_b200 as M` instead of `import model as M` -- and reproduces the
reference's float64 trajectory.  Nothing of GANTrainer / FusedAdamW / the repo's loss classes is involved."""
import os
import sys

import pytest
import torch

from conftest import ROOT, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("conv,pam,tol", [("fp32", "fp32", 1e-3), ("bf16x3", "fp16x3", 1e-2)])
def test_notebook_loop_on_the_repo_modules(golden, conv, pam, tol):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import notebook_step as NS
    import gan_danet_b200 as M
    from gan_danet_b200 import engine as E
    from gan_danet_b200.synthetic import make_batch
    g = golden("train_2steps_8x16")
    lr05, real, aux = make_batch(0, 2, 8, 16)
    torch.manual_seed(g["seed"])
    G = M.FlexibleUpsamplingModule(46)
    D = M.Discriminator1()
    G.apply(M.weights_init_normal)                      # GAN_DANet_train.ipynb:164-165
    D.apply(M.weights_init_normal)                      # the lazy fc1 is skipped, as the authors' torch did (SURVEY 8c caveat 3)
    D._materialise_fc1(real)                            # = the reference's first forward (consumes the RNG for nn.Linear's default init)
    torch.manual_seed(g["vgg_seed"])
    perc = M.PerceptualLoss(pretrained=False, device=torch.device("cpu"))
    G, D = G.to(DEV).train(), D.to(DEV).train()
    perc.vgg.to(DEV)
    perc.device = torch.device(DEV)
    G.set_pam_precision(pam)
    old = E.conv_precision
    E.set_conv_precision(conv)
    try:
        tr = NS.NotebookTrainer(M, G, D, epochs=g["epochs"], device=DEV, perceptual=perc)
        assert type(tr.optimizer_D) is torch.optim.AdamW and type(tr.adversarial_loss) is torch.nn.BCEWithLogitsLoss
        for step in range(2):
            out = tr.step(lr05, real, aux, epoch=g["epoch"])
            assert all(p.grad is not None for p in D.parameters())       # :260 D(hr_generated) left gradients on D, as in the notebook
            for k, v in g["history"][step].items():
                assert abs(out[k] - v) <= tol * max(abs(v), 1e-3), (step, k, out[k], v)
    finally:
        E.set_conv_precision(old)
    assert rel_err(G.final.weight, g["final_w"]) < 10 * tol
    assert rel_err(D.fc2.weight, g["d_fc2_w"]) < 10 * tol
    assert rel_err(G.initial[1].running_mean, g["initial_bn_rm"]) < 1e-3
