"""Drop-in for the reference's import shim ``model.py:1-5`` (``from model import FlexibleUpsamplingModule, ...``)."""
from .models import *  # noqa: F401,F403
from .models import __all__  # noqa: F401
