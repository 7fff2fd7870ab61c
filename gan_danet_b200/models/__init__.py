"""GAN-DANet model components on B200 kernels; same exports as the reference ``models/__init__.py:1-23``."""
from .generator import (
    CBAMBlock,
    FlexibleUpsamplingModule,
    OriginalRelationshipLearner,
    SqueezeExcitation,
)
from .discriminator import Discriminator1, SRGAND
from .losses import PerceptualLoss, SSIM, TVLoss
from .utils import weights_init_normal

__all__ = [
    "CBAMBlock",
    "FlexibleUpsamplingModule",
    "OriginalRelationshipLearner",
    "SqueezeExcitation",
    "Discriminator1",
    "SRGAND",
    "PerceptualLoss",
    "SSIM",
    "TVLoss",
    "weights_init_normal",
]
