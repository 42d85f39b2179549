"""Fused decode + objective evaluation kernels (validation path), one per problem.

All functions take/return CUDA fp32 tensors and run the hand-written kernels behind the
C-ABI (diffsg_minmax, diffsg_objective_msr, diffsg_decode_nu, diffsg_rate_nu,
diffsg_decode_co, diffsg_cost_co).  Reference semantics:
  MSR  ddpm_opt/classifier_free_MSR.py:239-245 (decoder), :284-288 (rate)
  NU   ddpm_opt/classifier_free_NU.py:267-276 (decoder), :279-303 (rate_calc)
  CO   ddpm_opt/classifier_free_CO.py:281-290 (decoder), :255-278 (cost_calc)
"""
from __future__ import annotations

import torch

from . import _lib


def _prep(t: torch.Tensor, cols=None) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.DiffsgError(f"objective kernels need CUDA tensors, got '{t.device}' (no CPU path)")
    t = t.detach().to(torch.float32)
    if t.dim() == 1:
        t = t[None, :]
    if cols is not None and t.shape[1] != cols:
        raise ValueError(f"expected {cols} columns, got {tuple(t.shape)}")
    return t.contiguous()


def _call(name, *args):
    lib = _lib.load()
    _lib.check(getattr(lib, name)(*args, _lib.stream_ptr()), name)


def global_minmax(y: torch.Tensor, col0=0, width=None) -> torch.Tensor:
    """(min, max) over y[:, col0:col0+width] as a 2-element CUDA tensor (no host sync)."""
    y = _prep(y)
    width = y.shape[1] - col0 if width is None else width
    mm = torch.empty(2, dtype=torch.float32, device=y.device)
    with torch.cuda.device(y.device):
        _call("diffsg_minmax", y.data_ptr(), y.shape[0], y.shape[1], col0, width, mm.data_ptr())
    return mm


def msr_decode_rate(y_pred, g, W=1.0, return_alloc=False):
    """p = W * softmax((y - min) / (max - min)), rate = sum_j log2(1 + p_j g_j).  g: raw gains."""
    y = _prep(y_pred)
    g = _prep(g, y.shape[1])
    B, M = y.shape
    mm = global_minmax(y)
    rate = torch.empty(B, dtype=torch.float32, device=y.device)
    p = torch.empty_like(y) if return_alloc else None
    with torch.cuda.device(y.device):
        _call("diffsg_objective_msr", y.data_ptr(), g.data_ptr(), mm.data_ptr(), float(W),
              p.data_ptr() if p is not None else None, rate.data_ptr(), B, M)
    return (rate, p) if return_alloc else rate


def msr_rate(p, g):
    """rate = sum_j log2(1 + p_j g_j) for a given allocation (labels)."""
    p = _prep(p)
    g = _prep(g, p.shape[1])
    rate = torch.empty(p.shape[0], dtype=torch.float32, device=p.device)
    with torch.cuda.device(p.device):
        _call("diffsg_rate_msr", p.data_ptr(), g.data_ptr(), rate.data_ptr(), p.shape[0], p.shape[1])
    return rate


def nu_decode(y_pred, width, height, P_sum):
    y = _prep(y_pred)
    B, K = y.shape[0], y.shape[1] - 2
    mm = global_minmax(y, 0, 2)
    dec = torch.empty_like(y)
    with torch.cuda.device(y.device):
        _call("diffsg_decode_nu", y.data_ptr(), mm.data_ptr(), float(width), float(height), float(P_sum),
              dec.data_ptr(), B, K)
    return dec


def nu_rate(decoded, xy):
    """NOMA-UAV sum rate; decoded: [B, 2+K] (uav x, y, powers), xy: [B, 2K] user coordinates."""
    d = _prep(decoded)
    K = d.shape[1] - 2
    x = _prep(xy, 2 * K)
    rate = torch.empty(d.shape[0], dtype=torch.float32, device=d.device)
    with torch.cuda.device(d.device):
        _call("diffsg_rate_nu", d.data_ptr(), x.data_ptr(), rate.data_ptr(), d.shape[0], K)
    return rate


def co_decode(y_pred):
    y = _prep(y_pred)
    dec = torch.empty_like(y)
    with torch.cuda.device(y.device):
        _call("diffsg_decode_co", y.data_ptr(), dec.data_ptr(), y.shape[0], y.shape[1])
    return dec


def co_cost(x, alloc):
    """Offloading cost; x: [B, 3n] (local, transition, ideal-exec per node), alloc: [B, n]."""
    a = _prep(alloc)
    x = _prep(x, 3 * a.shape[1])
    cost = torch.empty(a.shape[0], dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        _call("diffsg_cost_co", x.data_ptr(), a.data_ptr(), cost.data_ptr(), a.shape[0], a.shape[1])
    return cost
