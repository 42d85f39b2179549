"""Batched forwards of the reference's comparison baselines: the MTFNN multi-layer perceptrons
(baselines/MTFNN.py:43-52 CO, :122-131 MSR, :187-211 NU) and the PPO agent's actor / critic (baselines/PPO.py:33-70).

The modules keep the reference's parameter names and shapes, so `ckpts/mtfnn_{co,msr_3c,msr_80c,nu}.pt` and
`ckpts/ppo_{co,msr_3c,msr_80c,nu}.pt` strict-load; the forward of a whole batch is ONE launch of the library's small-MLP
kernel (`diffsg_mlp_forward`) instead of a chain of torch ops.  Like the rest of the package there is no CPU path.
Only inference is covered: these are baselines for side-by-side quality tables, not part of the DDPM hot path.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict

import torch
import torch.nn as nn

from . import _lib

ACT = {nn.ReLU: 1, nn.Tanh: 2, nn.Sigmoid: 3}


def _mlp_plan(layers):
    """[Linear, act?, Linear, act?, ..., head?] -> (linears, act codes, head code); Softmax only as the last module."""
    lin, acts, head = [], [], 0
    for m in layers:
        if isinstance(m, nn.Linear):
            lin.append(m)
            acts.append(0)
        elif type(m) in ACT:
            if not lin or acts[-1] != 0:
                raise _lib.DiffsgError("baselines: activation without a preceding Linear")
            acts[-1] = ACT[type(m)]
        elif isinstance(m, nn.Softmax):
            head = 1
        else:
            raise _lib.DiffsgError(f"baselines: unsupported module {type(m).__name__} in an MLP")
    if not lin:
        raise _lib.DiffsgError("baselines: no Linear layer")
    return lin, acts, head


def mlp_forward(layers, x: torch.Tensor, head: int | None = None, head_split: int = 0) -> torch.Tensor:
    """Run `layers` (an iterable of nn.Linear / nn.ReLU / nn.Tanh / nn.Sigmoid [/ final nn.Softmax]) on x [B, in] in one
    kernel launch.  `head` overrides the head found in `layers` (2 = sigmoid on the first `head_split` columns, softmax on
    the rest: the NU MTFNN)."""
    if not x.is_cuda:
        raise _lib.DiffsgError(f"diffsg_b200 runs on CUDA only: got a tensor on '{x.device}' (no CPU implementation)")
    lin, acts, found = _mlp_plan(list(layers))
    head = found if head is None else head
    x = x.detach().to(torch.float32).contiguous()
    B = x.shape[0]
    blob = torch.cat([t.detach().to(device=x.device, dtype=torch.float32).reshape(-1) for m in lin for t in (m.weight, m.bias)])
    out = torch.empty(B, lin[-1].out_features, dtype=torch.float32, device=x.device)
    if B == 0:
        return out
    n = len(lin)
    dims = (C.c_int32 * n)(*[m.out_features for m in lin])
    codes = (C.c_int32 * n)(*acts)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().diffsg_mlp_forward(x.data_ptr(), blob.data_ptr(), B, lin[0].in_features, n, dims, codes,
                                                  head, head_split, out.data_ptr(), _lib.stream_ptr()), "diffsg_mlp_forward")
    return out


class BatchedSequential(nn.Sequential):
    """nn.Sequential whose forward is the fused kernel (inference); same state_dict keys as the reference's
    `nn.Sequential(OrderedDict([('lin1', ...), ('act1', ...), ...]))` models."""

    def forward(self, x):
        return mlp_forward(self, x)


def mtfnn_co_model(in_dim, out_dim):
    """MTFNN for computation offloading (baselines/MTFNN.py:43-52): 32-64-16, Sigmoid head."""
    return BatchedSequential(OrderedDict([("lin1", nn.Linear(in_dim, 32)), ("act1", nn.ReLU()), ("lin2", nn.Linear(32, 64)),
                                          ("act2", nn.ReLU()), ("lin3", nn.Linear(64, 16)), ("act3", nn.ReLU()),
                                          ("lin4", nn.Linear(16, out_dim)), ("act4", nn.Sigmoid())]))


def mtfnn_msr_model(in_dim, out_dim):
    """MTFNN for the sum-rate problem (baselines/MTFNN.py:122-131): 8-16-8, Softmax head."""
    return BatchedSequential(OrderedDict([("lin1", nn.Linear(in_dim, 8)), ("act1", nn.ReLU()), ("lin2", nn.Linear(8, 16)),
                                          ("act2", nn.ReLU()), ("lin3", nn.Linear(16, 8)), ("act3", nn.ReLU()),
                                          ("lin4", nn.Linear(8, out_dim)), ("act4", nn.Softmax(dim=1))]))


class MTFNN(nn.Module):
    """MTFNN for NOMA-UAV (baselines/MTFNN.py:187-211): 64-32-16-32, sigmoid on the UAV position, softmax on the powers."""

    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.lin1, self.act1 = nn.Linear(in_dim, 64), nn.ReLU()
        self.lin2, self.act2 = nn.Linear(64, 32), nn.ReLU()
        self.lin3, self.act3 = nn.Linear(32, 16), nn.ReLU()
        self.lin4, self.act4 = nn.Linear(16, 32), nn.ReLU()
        self.lin5 = nn.Linear(32, out_dim)
        self.act51, self.act52 = nn.Sigmoid(), nn.Softmax(dim=1)

    def forward(self, x):
        return mlp_forward([self.lin1, self.act1, self.lin2, self.act2, self.lin3, self.act3, self.lin4, self.act4, self.lin5],
                           x, head=2, head_split=2)


class PPOAgent(nn.Module):
    """Actor / critic of the PPO baseline (baselines/PPO.py:33-70): `forward(state) -> (value, Normal(mu, exp(log_std)))`."""

    def __init__(self, state_dim, action_dim):
        super().__init__()
        self.state_dim, self.action_dim = state_dim, action_dim
        mk = lambda n_out: nn.Sequential(nn.Linear(state_dim, 64), nn.Tanh(), nn.Linear(64, 16), nn.Tanh(),
                                         nn.Linear(16, 32), nn.Tanh(), nn.Linear(32, n_out))
        self.critic = mk(1)
        self.actor = mk(action_dim)
        self.log_std = nn.Parameter(torch.zeros(1, action_dim))

    def forward(self, state):
        value = mlp_forward(self.critic, state)
        mu = mlp_forward(self.actor, state)
        return value, torch.distributions.Normal(mu, self.log_std.exp())
