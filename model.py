"""Top-level shim so the reference notebooks' ``from model import ...`` (GAN_DANet_train.ipynb:20, test.ipynb:34,
deep_ensemble.ipynb:46) resolves to the B200 implementation when this repo is on ``sys.path``."""
from gan_danet_b200.models import *  # noqa: F401,F403
from gan_danet_b200.models import __all__  # noqa: F401
