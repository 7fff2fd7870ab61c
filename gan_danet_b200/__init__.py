"""gan_danet_b200 -- the GAN-DANet generator/discriminator training step as hand-written sm_100a CUDA kernels behind
the reference's ``models`` module API.  ``from gan_danet_b200.model import *`` mirrors the reference's ``model.py``."""
from .models import (  # noqa: F401
    CBAMBlock,
    Discriminator1,
    FlexibleUpsamplingModule,
    OriginalRelationshipLearner,
    PerceptualLoss,
    SRGAND,
    SSIM,
    SqueezeExcitation,
    TVLoss,
    weights_init_normal,
)
from .models.losses import BCEWithLogitsLoss, L1Loss, MSELoss  # noqa: F401
from . import ops  # noqa: F401,E402  registers torch.ops.gandanet.* (CUDA dispatch key only; no GPU needed to register)

__version__ = "0.1.0"
