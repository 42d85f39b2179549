"""Deterministic stand-in weights for the configurations whose checkpoints are missing from
the reference repo.  TEST INFRASTRUCTURE ONLY.

Shared by oracle/make_golden.py (which feeds them to the unmodified reference) and the
tests (which feed them to the kernels), so both sides evaluate the same network without
shipping megabytes of weights.  Values come from a seeded torch CPU generator in sorted-key
order; magnitudes are those of a trained network rather than `init_weights` (std 0.01),
whose near-zero outputs would make parity checks vacuous."""
from __future__ import annotations

import torch

CONFIGS = {
    # name: (problem kind, UNet1D kwargs)      source of the hyper-parameters
    "msr3c": ("msr", dict(input_dim=3, proj_dim=128, cond_dim=3, dims=(64, 32, 16, 8),
                          is_attn=(False,) * 4, middle_attn=False, n_blocks=2)),   # MSR.py:262-263
    "msr80c": ("msr", dict(input_dim=80, proj_dim=128, cond_dim=80, dims=(64, 32, 16, 8),
                           is_attn=(False,) * 4, middle_attn=False, n_blocks=2)),  # ASSUMED
    "co": ("co", dict(input_dim=3, proj_dim=64, cond_dim=9, dims=(64, 32, 16, 8),
                      is_attn=(False,) * 4, middle_attn=False, n_blocks=3)),       # CO.py:307-308
    "nu_like": ("nu", dict(input_dim=5, proj_dim=32, cond_dim=6, dims=(32, 16, 8),
                           is_attn=(False,) * 3, middle_attn=False, n_blocks=2)),  # NU.py:322-323
    "attn": ("msr", dict(input_dim=4, proj_dim=16, cond_dim=5, dims=(16, 8),
                         is_attn=(True, False), middle_attn=True, n_blocks=1)),    # UNetCF.py:98-157
}


def make_state_dict(shapes: dict, seed: int = 1234) -> dict:
    """{key: shape} -> {key: fp32 tensor}. Linear weights ~ U(-1,1)*sqrt(3/fan_in) (unit-gain),
    LayerNorm gains ~ 1 + 0.2 N(0,1), every bias ~ 0.1 N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for key in sorted(shapes):
        shape = tuple(shapes[key])
        if len(shape) == 2:
            w = (torch.rand(shape, generator=g) * 2 - 1) * (3.0 / shape[1]) ** 0.5
        elif key.endswith("weight"):
            w = 1.0 + 0.2 * torch.randn(shape, generator=g)
        else:
            w = 0.1 * torch.randn(shape, generator=g)
        out[key] = w.to(torch.float32)
    return out
