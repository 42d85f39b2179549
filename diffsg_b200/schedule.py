"""Noise schedule + weight init helpers (reference ddpm_opt/diffusion.py:17-35, 82-84)."""
from __future__ import annotations

import numpy as np
import torch.nn as nn

BETA_CLIP = 0.84  # reference diffusion.py:34


def generate_cosine_schedule(T: int, s: float = 0.008) -> np.ndarray:
    """betas[T] (float64) of the cosine schedule, each clipped at 0.84."""
    t = np.arange(T + 1, dtype=np.float64)
    abar = np.cos((t / T + s) / (1 + s) * np.pi / 2) ** 2
    abar = abar / abar[0]
    return np.minimum(1.0 - abar[1:] / abar[:-1], BETA_CLIP)


def generate_linear_schedule(T: int, low: float, high: float) -> np.ndarray:
    return np.linspace(low, high, T)


def init_weights(m):
    """`module.apply(init_weights)`: every nn.Linear weight ~ N(0, 0.01^2); biases untouched."""
    if type(m) is nn.Linear:
        nn.init.normal_(m.weight, std=0.01)
