"""Synthetic "trained-like" weights and MPII-shaped inputs for benchmarks and demos (no checkpoint or
dataset is reachable offline).  Independent of oracle/: walks the module's own state_dict."""
from __future__ import annotations

import torch
import torch.nn as nn


def trained_like_init_(model: nn.Module, seed: int = 0) -> nn.Module:
    """He-normal convs, non-trivial BatchNorm affine / running statistics, small gamma on the last BatchNorm of
    every residual branch so activations stay O(1) through the 33 bottlenecks (also in fp16)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    with torch.no_grad():
        for name, m in model.named_modules():
            if isinstance(m, nn.Conv2d):
                fan_in = m.in_channels * m.kernel_size[0] * m.kernel_size[1]
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) * (2.0 / fan_in) ** 0.5)
                if m.bias is not None:
                    m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.05)
            elif isinstance(m, nn.BatchNorm2d):
                gamma = 0.75 + 0.5 * torch.rand(m.weight.shape, generator=g)
                if name.endswith(".bn3"):
                    gamma = gamma * 0.3
                m.weight.copy_(gamma)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(0.6 + 0.8 * torch.rand(m.running_var.shape, generator=g))
    return model


def mpii_like_input(n: int, h: int, w: int, seed: int = 0) -> torch.Tensor:
    """(uint8 - 128) / 256, the normalisation of utils/mpii_data.py:184-185."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1000 + seed)
    return (torch.randint(0, 256, (n, 3, h, w), generator=g).float() - 128.0) / 256.0
