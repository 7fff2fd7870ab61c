"""The reference's training loop, restated over ANY implementation of its module API.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as gan_danet_oracle.py).

``NotebookTrainer`` restates ``ModelTrainer.__init__`` / ``ModelTrainer.train`` of /root/reference/GAN_DANet_train.ipynb (cell "class ModelTrainer":
optimisers :182-183, losses :190-194, one iteration of the inner loop :225-269) with STOCK PyTorch pieces exactly as the notebook has them --
``torch.optim.AdamW``, ``torch.nn.BCEWithLogitsLoss()``, ``torch.nn.MSELoss()``, ``F.interpolate(..., mode='bicubic')`` for the input preparation,
``D(hr_generated)`` with D's parameters still requiring grad in the generator step -- parameterised by the namespace ``M`` that provides the
module classes (``FlexibleUpsamplingModule``, ``Discriminator1``, ``SSIM``, ``TVLoss``, ``PerceptualLoss``):

* ``M`` = the reference's own ``models`` package (vendored by oracle/build_ref.py into the git-ignored oracle/_ref/, or imported from
  /root/reference): the reference arm of bench.py (``cpu_baseline.kind = "reference"``) and the generator of golden vectors;
* ``M`` = ``gan_danet_b200`` on a CUDA device: the DROP-IN test -- the notebook's loop runs unchanged on this repo's modules
  (tests/test_gpu_dropin.py).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F


class NotebookTrainer:
    def __init__(self, M, G, D, *, epochs: int, device, perceptual=None, tv_weight: float = 1e-5):
        self.M, self.G, self.D, self.epochs, self.device = M, G, D, epochs, torch.device(device)
        # GAN_DANet_train.ipynb:182-183
        self.optimizer_D = torch.optim.AdamW(D.parameters(), lr=0.0004, betas=(0.5, 0.999), weight_decay=1e-4)
        self.optimizer_U = torch.optim.AdamW(G.parameters(), lr=0.0002, betas=(0.5, 0.999), weight_decay=1e-4)
        # :190-194 (PerceptualLoss(use_gpu=...) in the notebook downloads VGG19 weights; offline the caller passes a randomly initialised one)
        self.adversarial_loss = torch.nn.BCEWithLogitsLoss()
        self.pixelwise_loss = torch.nn.MSELoss()
        self.ssim_loss = M.SSIM(window_size=11, size_average=True).to(self.device)
        self.tv_loss = M.TVLoss(weight=tv_weight).to(self.device)
        self.perceptual_loss = perceptual

    def step(self, lr_grace_05: torch.Tensor, lr_grace_025: torch.Tensor, hr_aux: torch.Tensor, epoch: int) -> Dict[str, float]:
        """One iteration of ``for lr_grace_05, lr_grace_025, hr_aux in self.train_loader`` (:225-269).  Inputs are HOST tensors as the DataLoader yields
        them; the first interpolation runs on the host before ``.to(device)``, as in the notebook (:226-227)."""
        dev = self.device
        lr_grace = F.interpolate(lr_grace_05, scale_factor=0.5, mode="bicubic", align_corners=False)      # :226
        lr_grace, hr_aux = lr_grace.to(dev), hr_aux.to(dev)                                                # :227
        lr_grace_025 = lr_grace_025.to(dev)                                                                # :228
        downsampled_aux = F.interpolate(hr_aux, scale_factor=0.25, mode="bicubic", align_corners=False)   # :231
        combined_input = torch.cat([lr_grace, downsampled_aux], dim=1)                                     # :232
        hr_generated = self.G(combined_input)                                                              # :243

        self.optimizer_D.zero_grad()                                                                       # :246
        real_output = self.D(lr_grace_025)
        fake_output = self.D(hr_generated.detach())
        real_labels = torch.ones_like(real_output, device=dev)
        fake_labels = torch.zeros_like(fake_output, device=dev)
        loss_D = (self.adversarial_loss(real_output, real_labels) + self.adversarial_loss(fake_output, fake_labels)) / 2    # :252-254
        loss_D.backward()
        self.optimizer_D.step()                                                                            # :256

        self.optimizer_U.zero_grad()                                                                       # :259
        fake_output = self.D(hr_generated)                                                                 # :260 (D's parameters require grad here)
        loss_G_adv = self.adversarial_loss(fake_output, real_labels)
        loss_G_pixel = self.pixelwise_loss(hr_generated, lr_grace_025)
        loss_G_ssim = 1 - self.ssim_loss(hr_generated, lr_grace_025)                                       # :263, evaluated, never in the objective
        loss_G_tv = self.tv_loss(hr_generated)
        loss_weight = epoch / self.epochs                                                                  # :266
        loss_G = (1 - loss_weight) * loss_G_pixel + loss_weight * loss_G_adv + loss_G_tv
        loss_perc: Optional[torch.Tensor] = None
        if self.perceptual_loss is not None:
            loss_perc = self.perceptual_loss(hr_generated, lr_grace_025)
            loss_G = loss_G + loss_perc                                                                    # :267
        loss_G.backward()
        self.optimizer_U.step()                                                                            # :269
        out = {"loss_D": loss_D.item(), "loss_G": loss_G.item(), "adv": loss_G_adv.item(), "pixel": loss_G_pixel.item(),     # .item(): :271-272
               "ssim": float(loss_G_ssim), "tv": loss_G_tv.item()}
        if loss_perc is not None:
            out["perceptual"] = loss_perc.item()
        return out


def init_like_the_authors(M, G, D, sample_real: torch.Tensor) -> None:
    """``weights_init_normal`` on G and on D's convs + fc2; the lazy ``fc1`` keeps nn.Linear's default init from its first forward (SURVEY 8c caveat 3:
    torch >= 2.x raises when ``.apply`` meets an uninitialised LazyLinear, so the authors' effective init is this one)."""
    G.apply(M.weights_init_normal)
    for mod in (D.conv1, D.conv2, D.conv3, D.conv4, D.fc2):
        mod.apply(M.weights_init_normal)
    with torch.no_grad():
        D(sample_real)


def load_reference_models(ref_root: str):
    """Imports the reference's ``models`` package from ``ref_root`` (oracle/_ref or /root/reference) under its own name."""
    import importlib
    import sys
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
        del sys.modules[k]
    sys.path.insert(0, ref_root)
    try:
        return importlib.import_module("models")
    finally:
        sys.path.pop(0)
