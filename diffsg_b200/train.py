"""Differentiable UNet1D forward for training (eps-MSE step of DDPM.forward).

Every Linear of the training graph — forward, dgrad and wgrad — runs on this library's tcgen05 kernels
(`csrc/train_tc.cu`, C-ABI diffsg_tlin_forward / diffsg_tlin_backward: bf16 hi+lo operands, fp32 TMEM accumulators), with
the LayerNorm -> Swish in front of a Linear fused into its operand prologue (forward, wgrad) and into the dgrad
epilogue (backward), the bias / time-embedding / condition-embedding / residual adds fused into the forward epilogue,
and the bias gradient and the scatter of the gathered time term computed on the tensor cores inside wgrad.
PyTorch autograd is only the glue between the fused nodes: a ResidualBlock is three nodes (forward: one launch each,
backward: ONE launch each for its dgrad + wgrad problems), all blocks' time-embedding Linears are one node on the T
time rows (`time_table`), the 80-channel net is 94 forward + 94 backward launches.
Parameter gradients are accumulated IN PLACE into `p.grad` by the wgrad / dgrad kernels (`red.global.add`), so
autograd launches no accumulation kernels for the ~400 parameter tensors.

Semantics: reference ddpm_opt/UNetCF.py:83-95 (ResidualBlock), :123-157 (attention), :318-356 (UNet1D.forward);
loss: classifier_free_MSR.py:100-112.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn.functional as F

from . import _lib
from .packer import sinusoid
from .unet import AttentionBlock, DownBlock, UpBlock

MAX_GATHER_ROWS = 31        # one-hot columns the wgrad kernel appends next to the bias column (csrc/train_tc.cu)


def _ptr(t):
    return None if t is None else t.data_ptr()


def _mat(x0, x1=None):
    if x0 is None:
        return _lib.Mat(None, None, 0, 0)
    if x1 is None:
        return _lib.Mat(x0.data_ptr(), None, x0.shape[1], 0)
    return _lib.Mat(x0.data_ptr(), x1.data_ptr(), x0.shape[1], x1.shape[1])


def _prep(t):
    """fp32, contiguous, 2-D (or None)."""
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def _grad_target(p):
    """Where a parameter gradient is accumulated: `p.grad` itself for a leaf (allocated as zeros on first use, returned
    to autograd as None), a fresh zero tensor that is handed back to autograd otherwise."""
    if p.is_leaf:
        if p.grad is None:
            p.grad = torch.zeros_like(p, memory_format=torch.contiguous_format)
        if not p.grad.is_contiguous():
            raise _lib.DiffsgError("diffsg_b200 training needs contiguous .grad tensors")
        return p.grad, None
    g = torch.zeros_like(p, memory_format=torch.contiguous_format)
    return g, g


class _FusedLinear(torch.autograd.Function):
    """y = act(cat(x0, x1)) . w^T + b (+ cat(z0, z1) . w2^T + b2) (+ add) (+ gadd[gidx]);
    act = swish(LayerNorm(.; gamma, beta)) when gamma is given, identity otherwise."""

    @staticmethod
    def forward(ctx, x0, x1, gamma, beta, w, b, z0, z1, w2, b2, add, gadd, gidx, table=None):
        """`table = (column, holder)`: `gadd` is a wide [T, W] table (every block's time term, `time_table`) and this node
        adds its columns [column, column + N); the gradient goes straight into the same columns of `holder`'s buffer."""
        lib = _lib.load()
        if not x0.is_cuda:
            raise _lib.DiffsgError("diffsg_b200 training runs on CUDA only (no CPU implementation)")
        x0, x1, z0, z1, add, gadd = (_prep(t) for t in (x0, x1, z0, z1, add, gadd))
        col, holder = table if table is not None else (0, None)
        B, N = x0.shape[0], w.shape[0]
        K = x0.shape[1] + (x1.shape[1] if x1 is not None else 0)
        assert w.shape[1] == K and w.is_contiguous() and (b is None or b.is_contiguous())
        y = torch.empty(B, N, dtype=torch.float32, device=x0.device)
        mean = rstd = None
        if gamma is not None:
            mean = torch.empty(B, dtype=torch.float32, device=x0.device)
            rstd = torch.empty_like(mean)
        if gidx is not None:
            gidx = gidx.contiguous()
            assert gidx.dtype == torch.int64 and gidx.numel() == B
        a = _lib.TlinFwdArgs(a=_mat(x0, x1), w=w.data_ptr(), bias=_ptr(b), gamma=_ptr(gamma), beta=_ptr(beta),
                             mean=_ptr(mean), rstd=_ptr(rstd), a2=_mat(z0, z1), w2=_ptr(w2), bias2=_ptr(b2),
                             add=_ptr(add), gadd=None if gadd is None else gadd.data_ptr() + 4 * col, gidx=_ptr(gidx),
                             y=y.data_ptr(), B=B, N=N, gadd_ld=0 if gadd is None else gadd.shape[1])
        with torch.cuda.device(x0.device):
            _lib.check(lib.diffsg_tlin_forward(C.byref(a), _lib.stream_ptr()), "diffsg_tlin_forward")
        ctx.save_for_backward(x0, x1, z0, z1, mean, rstd, gidx)
        ctx.params = (gamma, beta, w, b, w2, b2)
        ctx.gadd_rows = 0 if gadd is None else gadd.shape[0]
        ctx.table = None if holder is None else (col, holder, gadd.shape[1], holder.claim())
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x0, x1, z0, z1, mean, rstd, gidx = ctx.saved_tensors
        gamma, beta, w, b, w2, b2 = ctx.params
        need = ctx.needs_input_grad
        dy = _prep(dy)
        B, N = dy.shape
        dev = dy.device
        out = [None] * 14

        def target(i, p):
            if p is None or not need[i]:
                return None
            t, ret = _grad_target(p)
            out[i] = ret
            return t

        wg, dg = [], []
        keep = []                                            # temporaries the kernels write into: alive until the launch
        # ---- wgrad of the main segment (+ bias gradient, + scatter of the gathered add)
        dw, db = target(4, w), target(5, b)
        dgadd, dgadd_ptr, dgadd_ld = None, None, 0
        if need[11] and ctx.gadd_rows:
            if ctx.table is not None:
                col, holder, width, returns_it = ctx.table
                dgadd = holder.buffer(ctx.gadd_rows, width, dev)
                dgadd_ptr, dgadd_ld = dgadd.data_ptr() + 4 * col, width
                out[11] = dgadd if returns_it else None      # the first consumer (last in backward) hands the table back
                if returns_it:
                    holder.release()
            else:
                dgadd = torch.zeros(ctx.gadd_rows, N, dtype=torch.float32, device=dev)
                dgadd_ptr = dgadd.data_ptr()
                out[11] = dgadd
        if dw is not None or db is not None or dgadd is not None:
            if dw is None:                                   # frozen weight: the kernel still needs a target
                dw = torch.zeros_like(w)
                keep.append(dw)
            wg.append(_lib.TlinWgradArgs(dy=dy.data_ptr(), a=_mat(x0, x1), gamma=_ptr(gamma), beta=_ptr(beta),
                                         mean=_ptr(mean), rstd=_ptr(rstd), gidx=_ptr(gidx) if dgadd is not None else None,
                                         dw=dw.data_ptr(), dbias=_ptr(db), dgadd=dgadd_ptr, B=B, N=N,
                                         gadd_rows=ctx.gadd_rows if dgadd is not None else 0, dgadd_ld=dgadd_ld))
        # ---- wgrad of the second (identity) segment
        if w2 is not None:
            dw2, db2 = target(8, w2), target(9, b2)
            if dw2 is not None or db2 is not None:
                if dw2 is None:
                    dw2 = torch.zeros_like(w2)
                    keep.append(dw2)
                wg.append(_lib.TlinWgradArgs(dy=dy.data_ptr(), a=_mat(z0, z1), dw=dw2.data_ptr(), dbias=_ptr(db2), B=B, N=N))
        # ---- dgrad of the main segment (through the LayerNorm -> Swish backward when there is one)
        if need[0] or (x1 is not None and need[1]) or (gamma is not None and (need[2] or need[3])):
            dx0 = torch.empty_like(x0)
            dx1 = torch.empty_like(x1) if x1 is not None else None
            dgm = dbt = None
            if gamma is not None:
                dgm, dbt = target(2, gamma), target(3, beta)
                dgm = dgm if dgm is not None else torch.zeros_like(gamma)
                dbt = dbt if dbt is not None else torch.zeros_like(beta)
                keep += [dgm, dbt]
            dg.append(_lib.TlinDgradArgs(dy=dy.data_ptr(), w=w.data_ptr(), x=_mat(x0, x1) if gamma is not None else _mat(None),
                                         gamma=_ptr(gamma), beta=_ptr(beta), mean=_ptr(mean), rstd=_ptr(rstd), dres=_mat(None),
                                         dx=_mat(dx0, dx1), dgamma=_ptr(dgm), dbeta=_ptr(dbt), B=B, N=N, K=w.shape[1]))
            out[0] = dx0 if need[0] else None
            out[1] = dx1 if (x1 is not None and need[1]) else None
            keep += [dx0, dx1]
        # ---- dgrad of the second segment
        if w2 is not None and (need[6] or (z1 is not None and need[7])):
            dz0 = torch.empty_like(z0)
            dz1 = torch.empty_like(z1) if z1 is not None else None
            dg.append(_lib.TlinDgradArgs(dy=dy.data_ptr(), w=w2.data_ptr(), x=_mat(None), dres=_mat(None), dx=_mat(dz0, dz1),
                                         B=B, N=N, K=w2.shape[1]))
            out[6] = dz0 if need[6] else None
            out[7] = dz1 if (z1 is not None and need[7]) else None
            keep += [dz0, dz1]
        if wg or dg:
            dga = (_lib.TlinDgradArgs * max(len(dg), 1))(*dg)
            wga = (_lib.TlinWgradArgs * max(len(wg), 1))(*wg)
            with torch.cuda.device(dev):                     # ONE launch for the whole backward of the node
                _lib.check(lib.diffsg_tlin_backward(dga, len(dg), wga, len(wg), _lib.stream_ptr()), "diffsg_tlin_backward")
        del keep
        if need[10]:
            out[10] = dy
        return tuple(out)


class _TableGrad:
    """Gradient of a time table: ONE zero-initialised [T, W] buffer that every consumer's wgrad adds its columns into.
    The consumer created first runs last in the backward (each block's gradient depends on all later blocks), so it is
    the one that returns the finished buffer to autograd; the others return None."""

    def __init__(self):
        self._buf = None
        self._claimed = False

    def claim(self):
        first, self._claimed = not self._claimed, True
        return first

    def buffer(self, rows, width, device):
        if self._buf is None:
            self._buf = torch.zeros(rows, width, dtype=torch.float32, device=device)
        return self._buf

    def release(self):
        """Called by the consumer that hands the finished buffer to autograd: a second backward through the same graph
        (retain_graph) starts from a fresh one."""
        buf, self._buf = self._buf, None
        return buf


class _CatParams(torch.autograd.Function):
    """torch.cat(parameters, 0) whose backward adds the slices straight into the parameters' .grad (one multi-tensor add)."""

    @staticmethod
    def forward(ctx, *ps):
        ctx.ps = ps
        return torch.cat(ps, 0)

    @staticmethod
    def backward(ctx, g):
        parts = g.split([p.shape[0] for p in ctx.ps], 0)
        outs, tgt, src = [], [], []
        for p, part, need in zip(ctx.ps, parts, ctx.needs_input_grad):
            if need and p.is_leaf:
                tgt.append(_grad_target(p)[0])
                src.append(part)
                outs.append(None)
            else:
                outs.append(part if need else None)
        if tgt:
            torch._foreach_add_(tgt, src)
        return tuple(outs)


def time_table(temb_act, blocks):
    """Every block's `time_emb` Linear of the (hoisted) time rows as ONE node: [T, 4P] x cat(W_i)^T -> [T, sum N_i].
    Returns (table, column offsets, gradient holder)."""
    w = _CatParams.apply(*[b.time_emb.weight for b in blocks])
    bias = _CatParams.apply(*[b.time_emb.bias for b in blocks])
    cols, c = [], 0
    for b in blocks:
        cols.append(c)
        c += b.time_emb.weight.shape[0]
    return fused_linear(temb_act, None, weight=w, bias=bias), cols, _TableGrad()


def fused_linear(x0, lin, x1=None, norm=None, seg2=None, add=None, gadd=None, gidx=None, weight=None, bias=None, table=None):
    """One fused node.  `lin` (nn.Linear) or explicit (`weight`, `bias`); `norm` (nn.LayerNorm) puts
    LayerNorm -> Swish in front; `seg2 = (z0, z1, lin2)` accumulates a second Linear of another input."""
    w = lin.weight if weight is None else weight
    b = (lin.bias if lin is not None else None) if bias is None else bias
    z0 = z1 = w2 = b2 = None
    if seg2 is not None:
        z0, z1, l2 = seg2
        w2, b2 = l2.weight, l2.bias
    return _FusedLinear.apply(x0, x1, None if norm is None else norm.weight, None if norm is None else norm.bias,
                              w, b, z0, z1, w2, b2, add, gadd, gidx, table)


def ln_swish_linear(x, norm, lin):
    """lin(swish(norm(x))) as one tensor-core node."""
    return fused_linear(x, lin, norm=norm)


def _res(blk, x0, x1, temb_act, cond_act, gidx, tt=None):
    """ResidualBlock (UNetCF.py:83-95) as three fused nodes.  The time term comes from the shared time table `tt`
    (gather mode) or from the block's own time-embedding node."""
    if tt is not None:
        table, cols, holder = tt
        h = fused_linear(x0, blk.lin1, x1=x1, norm=blk.norm1, gadd=table, gidx=gidx, table=(cols[id(blk)], holder))
    elif gidx is not None:
        h = fused_linear(x0, blk.lin1, x1=x1, norm=blk.norm1, gadd=fused_linear(temb_act, blk.time_emb), gidx=gidx)
    else:
        h = fused_linear(x0, blk.lin1, x1=x1, norm=blk.norm1, add=fused_linear(temb_act, blk.time_emb))
    h = fused_linear(h, blk.lin2, norm=blk.norm2, seg2=(cond_act, None, blk.cond_emb))
    if isinstance(blk.shortcut, torch.nn.Linear):
        return fused_linear(h, blk.lin3, norm=blk.norm3, seg2=(x0, x1, blk.shortcut))
    assert x1 is None
    return fused_linear(h, blk.lin3, norm=blk.norm3, add=x0)


def _attn(att, x):
    if not isinstance(att, AttentionBlock):
        return x
    dk, nh = att.d_k, att.n_heads
    rows = torch.cat([torch.arange(h * 3 * dk + 2 * dk, (h + 1) * 3 * dk, device=x.device) for h in range(nh)])
    # sequence length 1: softmax == 1, the block is output(V(x)) + x
    v = fused_linear(x, None, weight=att.projection.weight[rows].contiguous(), bias=att.projection.bias[rows].contiguous())
    return fused_linear(v, att.output, add=x)


def unet_forward_train(model, x, t, cond, cond_mask, t_index=None, n_steps=None):
    """`t_index` [B] (integer step of every row) with `n_steps` = T hoists the time path exactly as the sampler
    does: TimeEmbedding and every block's `time_emb` Linear are evaluated on the T grid values i / T only
    ([T, 4P] instead of [B, 4P] operands: 60 % of the forward MACs of the 80c net disappear); the forward kernel
    gathers the row's time term in its epilogue and wgrad scatters its gradient back (one-hot columns on the tensor
    cores).  Without it `t` [1, B] may hold arbitrary values."""
    if not x.is_cuda:
        raise _lib.DiffsgError("diffsg_b200 training runs on CUDA only (no CPU implementation)")
    x = x.to(torch.float32).reshape(-1, model.input_dim)
    B = x.shape[0]
    te = model.time_emb
    gidx = None
    if t_index is not None:
        gidx = t_index.reshape(-1).to(torch.long)
        t = torch.arange(int(n_steps), device=x.device, dtype=torch.float32) / float(n_steps)
    e = sinusoid(t.reshape(-1).to(torch.float32), model.proj_dim)
    temb = fused_linear(F.silu(fused_linear(e, te.lin1)), te.lin2)
    if gidx is None and temb.shape[0] == 1 and B > 1:
        gidx = torch.zeros(B, dtype=torch.long, device=x.device)
    if gidx is not None and temb.shape[0] > MAX_GATHER_ROWS:
        temb, gidx = torch.index_select(temb, 0, gidx), None          # long schedules: gather once, add row by row
    temb_act = F.silu(temb)
    cond_act = F.silu(cond.to(torch.float32).reshape(B, -1) * cond_mask.to(torch.float32).reshape(-1, 1))
    tt = None
    if gidx is not None:
        # gather mode: all blocks' time-embedding Linears as one [T, 4P] x [4P, sum N] node, forward and backward
        blocks = ([m.res for m in model.down if isinstance(m, DownBlock)] + [model.middle.res1, model.middle.res2]
                  + [m.res for m in model.up if isinstance(m, UpBlock)])
        if all(p.is_leaf for b in blocks for p in b.time_emb.parameters()):
            table, cols, holder = time_table(temb_act, blocks)
            tt = (table, {id(b): c for b, c in zip(blocks, cols)}, holder)
    h = fused_linear(x, model.feature_proj)
    skips = [h]
    for m in model.down:
        if isinstance(m, DownBlock):
            h = _attn(m.attn, _res(m.res, h, None, temb_act, cond_act, gidx, tt))
        else:
            h = fused_linear(h, m.lin)
        skips.append(h)
    h = _res(model.middle.res1, h, None, temb_act, cond_act, gidx, tt)
    h = _attn(model.middle.attn, h)
    h = _res(model.middle.res2, h, None, temb_act, cond_act, gidx, tt)
    for m in model.up:
        if isinstance(m, UpBlock):
            h = _attn(m.attn, _res(m.res, h, skips.pop(), temb_act, cond_act, gidx, tt))      # cat(h, skip) is never formed
        else:
            h = fused_linear(h, m.lin)
    return fused_linear(h, model.final, norm=model.norm)
