"""Kernel-backed conditional vector U-Net (`UNet1D`).

Drop-in for the reference denoiser (reference `ddpm_opt/UNetCF.py:260-356`): same
constructor arguments, same parameter names / registration order (so reference
checkpoints strict-load), same `forward(x, t, cond, cond_mask)` signature.

The modules below are *parameter containers*: none of them evaluates the network with
torch ops.  `UNet1D.forward` lowers the module tree to a flat op program
(`diffsg_b200.packer`), hoists the batch-invariant time path into a bias table and runs
the hand-written sm_100a kernels behind the C-ABI (`include/diffsg_b200.h`).  There is
no CPU implementation; non-CUDA inputs raise.
"""
from __future__ import annotations

import torch
import torch.nn as nn


class TimeEmbedding(nn.Module):
    """Sinusoid -> Linear(P, 4P) -> Swish -> Linear(4P, 4P); `in_dim` = 4P.

    Parameter layout of reference `UNetCF.py:17-46` (`lin1`, `lin2`).
    """

    def __init__(self, in_dim: int):
        super().__init__()
        self.in_dim = in_dim
        self.lin1 = nn.Linear(in_dim // 4, in_dim)
        self.lin2 = nn.Linear(in_dim, in_dim)


class ResidualBlock(nn.Module):
    """3 x (LayerNorm -> Swish -> Linear) with time / condition biases and a shortcut.

    Parameter layout of reference `UNetCF.py:49-95`; `shortcut` is a Linear only when
    `in_dim != out_dim` (an `nn.Identity` otherwise, which owns no state_dict keys).
    """

    def __init__(self, in_dim: int, out_dim: int, time_dim: int, cond_dim: int):
        super().__init__()
        self.in_dim, self.out_dim = in_dim, out_dim
        self.norm1 = nn.LayerNorm(in_dim)
        self.lin1 = nn.Linear(in_dim, out_dim)
        self.norm2 = nn.LayerNorm(out_dim)
        self.lin2 = nn.Linear(out_dim, out_dim)
        self.norm3 = nn.LayerNorm(out_dim)
        self.lin3 = nn.Linear(out_dim, out_dim)
        self.shortcut = nn.Linear(in_dim, out_dim) if in_dim != out_dim else nn.Identity()
        self.time_emb = nn.Linear(time_dim, out_dim)
        self.cond_emb = nn.Linear(cond_dim, out_dim)


class AttentionBlock(nn.Module):
    """Single-token attention (reference `UNetCF.py:98-157`).

    The reference reshapes each vector to a length-1 sequence, so softmax == 1 and the
    block is exactly `x + output(V(x))` with `V` the last third of `projection`; `norm`
    is registered but never applied.  All three parameter groups are kept for
    state_dict compatibility.
    """

    def __init__(self, in_dim: int, n_heads: int = 1, d_k: int | None = None):
        super().__init__()
        d_k = in_dim if d_k is None else d_k
        self.n_heads, self.d_k = n_heads, d_k
        self.norm = nn.LayerNorm(in_dim)
        self.projection = nn.Linear(in_dim, n_heads * d_k * 3)
        self.output = nn.Linear(n_heads * d_k, in_dim)


class _ResStage(nn.Module):
    """`res` (+ optional `attn`) — shared shape of the reference Down/UpBlock."""

    def __init__(self, in_dim, out_dim, time_dim, cond_dim, has_attn):
        super().__init__()
        self.res = ResidualBlock(in_dim, out_dim, time_dim, cond_dim)
        self.attn = AttentionBlock(out_dim) if has_attn else nn.Identity()


class DownBlock(_ResStage):
    """reference `UNetCF.py:160-179`."""


class UpBlock(_ResStage):
    """reference `UNetCF.py:182-203`: consumes `cat(x, skip)` of width in+out."""

    def __init__(self, in_dim, out_dim, time_dim, cond_dim, has_attn):
        super().__init__(in_dim + out_dim, out_dim, time_dim, cond_dim, has_attn)


class MiddleBlock(nn.Module):
    """reference `UNetCF.py:206-227`."""

    def __init__(self, in_dim, time_dim, cond_dim, has_attn):
        super().__init__()
        self.res1 = ResidualBlock(in_dim, in_dim, time_dim, cond_dim)
        self.attn = AttentionBlock(in_dim) if has_attn else nn.Identity()
        self.res2 = ResidualBlock(in_dim, in_dim, time_dim, cond_dim)


class _Resample(nn.Module):
    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.lin = nn.Linear(in_dim, out_dim)


class Downsample(_Resample):
    """reference `UNetCF.py:245-257` (one Linear)."""


class Upsample(_Resample):
    """reference `UNetCF.py:230-242` (one Linear)."""


class UNet1D(nn.Module):
    """Vector U-Net denoiser; constructor mirrors reference `UNetCF.py:262-316`."""

    def __init__(self, input_dim=3, proj_dim=16, cond_dim=4, dims=(8, 4, 2),
                 is_attn=(False, False, False), middle_attn=False, n_blocks=2):
        super().__init__()
        self.input_dim, self.proj_dim, self.cond_dim = input_dim, proj_dim, cond_dim
        self.dims, self.is_attn = tuple(dims), tuple(is_attn)
        self.middle_attn, self.n_blocks = middle_attn, n_blocks
        tdim = proj_dim * 4
        levels = len(self.dims)

        self.feature_proj = nn.Linear(input_dim, proj_dim)
        self.time_emb = TimeEmbedding(tdim)

        down, width = [], proj_dim
        for lvl, nxt in enumerate(self.dims):
            down += [DownBlock(width, width, tdim, cond_dim, self.is_attn[lvl]) for _ in range(n_blocks)]
            down.append(Downsample(width, nxt))
            width = nxt
        down += [DownBlock(width, width, tdim, cond_dim, self.is_attn[-1]) for _ in range(n_blocks)]
        self.down = nn.ModuleList(down)

        self.middle = MiddleBlock(width, tdim, cond_dim, middle_attn)

        up = []
        for lvl in reversed(range(levels)):
            up += [UpBlock(width, width, tdim, cond_dim, self.is_attn[lvl]) for _ in range(n_blocks + 1)]
            nxt = self.dims[lvl - 1] if lvl > 0 else proj_dim
            up.append(Upsample(width, nxt))
            width = nxt
        up += [UpBlock(width, width, tdim, cond_dim, self.is_attn[0]) for _ in range(n_blocks + 1)]
        self.up = nn.ModuleList(up)

        self.norm = nn.LayerNorm(width)
        self.final = nn.Linear(width, input_dim)

        self._engine = None  # lazily-built kernel plan (diffsg_b200.engine.UNetEngine)
        self.precision = "auto"  # "auto" | "fp32" | "fp16x2" | "fp16x3" (see engine.UNetEngine)
        self._param_epoch = 0    # bumped by mark_params_changed()

    # ------------------------------------------------------------------ kernel glue
    def engine(self):
        """The kernel plan bound to this module's current parameters (built on demand)."""
        from .engine import UNetEngine
        if self._engine is None or self._engine.precision_request != self.precision:
            self._engine = UNetEngine(self, self.precision)
            self._engine.precision_request = self.precision
        return self._engine

    def mark_params_changed(self):
        """Tell the kernel plan that parameter VALUES changed through a path autograd's version
        counters do not see (e.g. an optimiser stepping a flat buffer the parameters are views of)."""
        self._param_epoch += 1

    def _apply(self, fn, *a, **k):  # .to()/.cuda()/.float(): packed weights are stale
        self._engine = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._engine = None
        return super().load_state_dict(*a, **k)

    def __deepcopy__(self, memo):  # EMA deep-copies the model; plans are not copyable
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for key, val in self.__dict__.items():
            new.__dict__[key] = None if key == "_engine" else copy.deepcopy(val, memo)
        return new

    def forward(self, x, t, cond, cond_mask):
        """eps = UNet(x[B,M], t[1,B] in [0,1), cond[B,C], cond_mask[B,1]).

        Semantics of reference `UNetCF.py:318-356`.  Runs the fused CUDA forward; under
        `torch.enable_grad()` with parameters requiring grad it is differentiable
        (`diffsg_b200.train`).
        """
        from .engine import unet_forward
        return unet_forward(self, x, t, cond, cond_mask)


def _forward_steps(self, x, ts, n_steps, cond, cond_mask):
    """`forward(x, ts / n_steps, cond, cond_mask)` for INTEGER steps `ts` [1, B] or [B]: the training graph
    evaluates the time path on the n_steps grid values only (diffsg_b200.train); inference is unchanged."""
    from .engine import unet_forward
    needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
    if needs_grad and x.is_cuda:
        from .train import unet_forward_train
        # the gather of the row's time term is fused into the forward epilogue and its scatter into wgrad: always hoist
        return unet_forward_train(self, x, None, cond, cond_mask, t_index=ts, n_steps=n_steps)
    if x.is_cuda:
        return self.engine().forward(x, None, cond, cond_mask, t_index=ts, n_steps=n_steps)   # cached step table, no host sync
    return unet_forward(self, x, ts / n_steps, cond, cond_mask)


UNet1D.forward_steps = _forward_steps


def infer_config_from_state_dict(sd, prefix="model."):
    """Recover UNet1D constructor arguments from a reference checkpoint's tensor shapes.

    The 80-channel hyper-parameters are not recorded in the reference repo,
    so a real `ddpm_msr_80c.pt` is loaded by reading its shapes.
    """
    def shape(name):
        return tuple(sd[prefix + name].shape)

    proj_dim, input_dim = shape("feature_proj.weight")
    cond_dim = None
    down_idx = sorted({int(k[len(prefix) + 5:].split(".")[0]) for k in sd if k.startswith(prefix + "down.")})
    dims, is_attn, n_blocks, run, run_attn = [], [], None, 0, False
    for i in down_idx:
        if prefix + f"down.{i}.lin.weight" in sd:
            dims.append(shape(f"down.{i}.lin.weight")[0])
            is_attn.append(run_attn)
            n_blocks = run if n_blocks is None else n_blocks
            run, run_attn = 0, False
        else:
            run += 1
            run_attn = run_attn or (prefix + f"down.{i}.attn.output.weight" in sd)
            if cond_dim is None:
                cond_dim = shape(f"down.{i}.res.cond_emb.weight")[1]
    middle_attn = prefix + "middle.attn.output.weight" in sd
    return dict(input_dim=input_dim, proj_dim=proj_dim, cond_dim=cond_dim, dims=tuple(dims),
                is_attn=tuple(is_attn), middle_attn=middle_attn, n_blocks=n_blocks)
