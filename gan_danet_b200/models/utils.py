"""Utility helpers for GAN-DANet models (reference ``models/utils.py:7-21``)."""
from __future__ import annotations

from torch import nn
from torch.nn.parameter import UninitializedParameter


def weights_init_normal(module: nn.Module) -> None:
    """Initialize common layers with Kaiming/Xavier schemes (same order and RNG consumption as the reference).

    A still-lazy ``nn.LazyLinear`` (``Discriminator1.fc1`` before its first forward, GAN_DANet_train.ipynb:164) is
    skipped: on the authors' torch that call was a no-op warning, on torch >= 2.11 it raises."""
    if isinstance(module, nn.Conv2d):
        nn.init.kaiming_normal_(module.weight, mode="fan_in", nonlinearity="relu")
        if module.bias is not None:
            nn.init.constant_(module.bias, 0)
    elif isinstance(module, nn.BatchNorm2d):
        nn.init.constant_(module.weight, 1)
        nn.init.constant_(module.bias, 0)
    elif isinstance(module, nn.Linear):
        if isinstance(module.weight, UninitializedParameter):
            return
        nn.init.xavier_normal_(module.weight)
        if module.bias is not None:
            nn.init.constant_(module.bias, 0)
    elif isinstance(module, nn.Parameter):  # pragma: no cover - dead branch kept from the reference
        nn.init.constant_(module, 0)


__all__ = ["weights_init_normal"]
