"""Lower a `UNet1D` module tree to the flat op program + parameter blob the kernels run.

What is decided here (host side, once per parameter version):

* topology -> `Op` list over four per-row buffers + a skip stack (include/diffsg_b200.h);
* weights  -> one fp32 blob: every nn.Linear stored transposed `[K][ldw]` (ldw = N rounded
  up to 4, zero padded) so a warp reads consecutive output columns; LayerNorm gamma/beta
  and biases inline;
* hoisting:
    - the whole time path (sinusoid -> TimeEmbedding MLP -> per-block `time_emb`) depends
      only on `t`, so it becomes a `[n_t, sum(out_dim)]` table added as a per-row bias;
    - `cond_emb.bias` is folded into `lin2.bias`; the unconditional pass (cond * 0 ->
      Swish(0) = 0) then needs no cond GEMM at all.
* the single-token attention block collapses to two Linears.

Reference semantics: ddpm_opt/UNetCF.py:30-46 (time), :83-95 (ResidualBlock), :123-157
(attention), :318-356 (forward).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch
import torch.nn as nn

from . import _lib
from .unet import (AttentionBlock, DownBlock, Downsample, MiddleBlock, ResidualBlock, UNet1D, UpBlock,
                   Upsample)

BUF_X, BUF_A, BUF_H, BUF_S = 0, 1, 2, 3


def _r4(n: int) -> int:
    return (n + 3) & ~3


@dataclass
class Program:
    ops: list = field(default_factory=list)          # list[dict] (Op fields)
    pieces: list = field(default_factory=list)       # list[(offset, callable -> flat fp32 tensor)]
    n_params: int = 0
    skip_widths: list = field(default_factory=list)
    time_blocks: list = field(default_factory=list)  # list[(t_off, nn.Linear)] per ResidualBlock
    tt_stride: int = 0
    in_buf: int = BUF_H
    out_buf: int = BUF_H
    max_width: int = 0
    input_dim: int = 0
    cond_dim: int = 0

    # ---- blob allocation
    def _alloc(self, n: int, make) -> int:
        off = self.n_params
        self.pieces.append((off, n, make))
        self.n_params += _r4(n)
        return off

    def add_linear(self, weight_fn, bias_fn, K: int, N: int):
        """weight_fn() -> [N, K] tensor (nn.Linear layout); stored as W^T [K][ldw]."""
        ldw = _r4(N)

        def make_w():
            w = weight_fn().detach().to(torch.float32)
            out = w.new_zeros(K, ldw)
            out[:, :N] = w.t()
            return out.reshape(-1)

        w_off = self._alloc(K * ldw, make_w)
        b_off = -1
        if bias_fn is not None:
            b_off = self._alloc(N, lambda: bias_fn().detach().to(torch.float32).reshape(-1))
        return w_off, b_off, ldw

    def add_vector(self, fn, n: int) -> int:
        return self._alloc(n, lambda: fn().detach().to(torch.float32).reshape(-1))

    # ---- op emission
    def gemm(self, src, dst, K, N, w_off, b_off, ldw, flags=0, t_off=-1, dcol=0):
        if b_off < 0:
            flags |= _lib.F_NOBIAS
            b_off = 0
        self.ops.append(dict(kind=_lib.OP_GEMM, src=src, dst=dst, K=K, N=N, flags=flags, w_off=w_off,
                             b_off=b_off, t_off=t_off, dcol=dcol, ldw=ldw))
        self.max_width = max(self.max_width, K, dcol + N)

    def lnsw(self, src, dst, D, ln: nn.LayerNorm):
        g = self.add_vector(lambda: ln.weight, D)
        b = self.add_vector(lambda: ln.bias, D)
        self.ops.append(dict(kind=_lib.OP_LNSW, src=src, dst=dst, K=0, N=D, flags=0, w_off=g, b_off=b,
                             t_off=-1, dcol=0, ldw=0))
        self.max_width = max(self.max_width, D)

    def push(self, src, D) -> int:
        slot = len(self.skip_widths)
        self.skip_widths.append(D)
        self.ops.append(dict(kind=_lib.OP_PUSH, src=src, dst=0, K=0, N=D, flags=0, w_off=0, b_off=0,
                             t_off=-1, dcol=slot, ldw=0))
        return slot

    def pop(self, slot, dst, D, dcol):
        self.ops.append(dict(kind=_lib.OP_POP, src=0, dst=dst, K=slot, N=D, flags=0, w_off=0, b_off=0,
                             t_off=-1, dcol=dcol, ldw=0))
        self.max_width = max(self.max_width, dcol + D)

    # ---- ctypes views
    def op_array(self):
        arr = (_lib.Op * len(self.ops))()
        for i, o in enumerate(self.ops):
            for k, v in o.items():
                setattr(arr[i], k, v)
        return arr

    def gemm_macs(self, include_cond=True):
        """Algorithmic multiply-accumulates per row-forward (un-padded shapes)."""
        x = sum(o["K"] * o["N"] for o in self.ops if o["kind"] == _lib.OP_GEMM and o["src"] != _lib.BUF_COND)
        c = sum(o["K"] * o["N"] for o in self.ops if o["kind"] == _lib.OP_GEMM and o["src"] == _lib.BUF_COND)
        return (x, c) if include_cond else x


def _lower_res(p: Program, blk: ResidualBlock, cur: int, alt: int) -> tuple[int, int]:
    """Emit one ResidualBlock; input of width blk.in_dim sits in `cur`. Returns (cur, alt)."""
    din, dout = blk.in_dim, blk.out_dim
    has_sc = isinstance(blk.shortcut, nn.Linear)
    t_off = p.tt_stride
    p.time_blocks.append((t_off, blk.time_emb))
    p.tt_stride += _r4(dout)

    p.lnsw(cur, BUF_A, din, blk.norm1)
    if has_sc:
        w, b, ld = p.add_linear(lambda: blk.shortcut.weight, lambda: blk.shortcut.bias, din, dout)
        p.gemm(cur, alt, din, dout, w, b, ld)
    w, b, ld = p.add_linear(lambda: blk.lin1.weight, lambda: blk.lin1.bias, din, dout)
    p.gemm(BUF_A, BUF_H, din, dout, w, b, ld, flags=_lib.F_TIME, t_off=t_off)
    p.lnsw(BUF_H, BUF_A, dout, blk.norm2)
    # lin2 bias absorbs cond_emb.bias (present in both the cond and the uncond pass)
    w, b, ld = p.add_linear(lambda: blk.lin2.weight, lambda: blk.lin2.bias + blk.cond_emb.bias, dout, dout)
    p.gemm(BUF_A, BUF_H, dout, dout, w, b, ld)
    cdim = blk.cond_emb.in_features
    w, _, ld = p.add_linear(lambda: blk.cond_emb.weight, None, cdim, dout)
    p.gemm(_lib.BUF_COND, BUF_H, cdim, dout, w, -1, ld, flags=_lib.F_ACC)
    p.lnsw(BUF_H, BUF_A, dout, blk.norm3)
    w, b, ld = p.add_linear(lambda: blk.lin3.weight, lambda: blk.lin3.bias, dout, dout)
    if has_sc:
        p.gemm(BUF_A, alt, dout, dout, w, b, ld, flags=_lib.F_ACC)
        return alt, cur
    p.gemm(BUF_A, cur, dout, dout, w, b, ld, flags=_lib.F_ACC)
    return cur, alt


def _lower_attn(p: Program, att, cur: int, D: int):
    """Single-token attention == x + output(V(x)); `norm` is unused by the reference."""
    if not isinstance(att, AttentionBlock):
        return
    dk, nh = att.d_k, att.n_heads
    rows = torch.cat([torch.arange(h * 3 * dk + 2 * dk, (h + 1) * 3 * dk) for h in range(nh)])
    w, b, ld = p.add_linear(lambda: att.projection.weight[rows], lambda: att.projection.bias[rows], D, nh * dk)
    p.gemm(cur, BUF_H, D, nh * dk, w, b, ld)
    w, b, ld = p.add_linear(lambda: att.output.weight, lambda: att.output.bias, nh * dk, D)
    p.gemm(BUF_H, cur, nh * dk, D, w, b, ld, flags=_lib.F_ACC)


def lower(model: UNet1D) -> Program:
    """Topology pass (no tensor data is touched; works for modules on any device)."""
    p = Program(input_dim=model.input_dim, cond_dim=model.cond_dim)
    cur, alt = BUF_X, BUF_S
    fp = model.feature_proj
    w, b, ld = p.add_linear(lambda: fp.weight, lambda: fp.bias, model.input_dim, model.proj_dim)
    p.gemm(BUF_H, cur, model.input_dim, model.proj_dim, w, b, ld)
    width = model.proj_dim
    stack = [(p.push(cur, width), width)]
    for m in model.down:
        if isinstance(m, DownBlock):
            cur, alt = _lower_res(p, m.res, cur, alt)
            width = m.res.out_dim
            _lower_attn(p, m.attn, cur, width)
        elif isinstance(m, Downsample):
            lin = m.lin
            w, b, ld = p.add_linear(lambda lin=lin: lin.weight, lambda lin=lin: lin.bias, lin.in_features, lin.out_features)
            p.gemm(cur, alt, lin.in_features, lin.out_features, w, b, ld)
            cur, alt = alt, cur
            width = lin.out_features
        else:
            raise TypeError(type(m))
        stack.append((p.push(cur, width), width))
    mid: MiddleBlock = model.middle
    cur, alt = _lower_res(p, mid.res1, cur, alt)
    _lower_attn(p, mid.attn, cur, width)
    cur, alt = _lower_res(p, mid.res2, cur, alt)
    for m in model.up:
        if isinstance(m, Upsample):
            lin = m.lin
            w, b, ld = p.add_linear(lambda lin=lin: lin.weight, lambda lin=lin: lin.bias, lin.in_features, lin.out_features)
            p.gemm(cur, alt, lin.in_features, lin.out_features, w, b, ld)
            cur, alt = alt, cur
            width = lin.out_features
        elif isinstance(m, UpBlock):
            slot, sw = stack.pop()
            p.pop(slot, cur, sw, width)
            assert m.res.in_dim == width + sw, (m.res.in_dim, width, sw)
            cur, alt = _lower_res(p, m.res, cur, alt)
            width = m.res.out_dim
            _lower_attn(p, m.attn, cur, width)
        else:
            raise TypeError(type(m))
    p.lnsw(cur, BUF_A, width, model.norm)
    fin = model.final
    w, b, ld = p.add_linear(lambda: fin.weight, lambda: fin.bias, width, model.input_dim)
    p.gemm(BUF_A, BUF_H, width, model.input_dim, w, b, ld)
    p.in_buf, p.out_buf = BUF_H, BUF_H
    return p


def pack_params(p: Program, device) -> torch.Tensor:
    """Materialise the parameter blob (fp32, on `device`) from the module's current values."""
    with torch.no_grad():
        blob = torch.zeros(max(p.n_params, 4), dtype=torch.float32, device=device)
        for off, n, make in p.pieces:
            blob[off:off + n] = make().to(device)
    return blob


def sinusoid(t: torch.Tensor, proj_dim: int) -> torch.Tensor:
    """[n] time values -> [n, proj_dim] features (reference UNetCF.py:35-40)."""
    half = proj_dim // 2
    scale = math.log(10_000) / (half - 1)
    freq = torch.exp(torch.arange(half, device=t.device) * -scale)
    ang = t.reshape(-1, 1).to(torch.float32) * freq[None, :]
    return torch.cat((ang.sin(), ang.cos()), dim=1)


def _swish(x):
    return x * torch.sigmoid(x)


def time_table(model: UNet1D, p: Program, t_values: torch.Tensor) -> torch.Tensor:
    """Hoisted time path: row r = concat_k time_emb_k(Swish(TimeEmbedding(t_values[r]))).

    Shape [len(t_values), p.tt_stride] fp32; column block k starts at the t_off recorded for
    ResidualBlock k.  (TimeEmbedding: UNetCF.py:30-46; per-block projection: UNetCF.py:91.)
    """
    te = model.time_emb
    with torch.no_grad():
        e = sinusoid(t_values, model.proj_dim)
        e = torch.nn.functional.linear(_swish(torch.nn.functional.linear(e, te.lin1.weight, te.lin1.bias)),
                                       te.lin2.weight, te.lin2.bias)
        a = _swish(e)
        tab = torch.zeros(a.shape[0], max(p.tt_stride, 4), dtype=torch.float32, device=a.device)
        for t_off, lin in p.time_blocks:
            tab[:, t_off:t_off + lin.out_features] = torch.nn.functional.linear(a, lin.weight, lin.bias)
    return tab.contiguous()
