"""Lower a `UNet1D` to the tensor-core engine's stage program (include/diffsg_b200.h,
"tensor-core program").

One CTA carries a 128-row tile; TMEM lane == row.  The network becomes a list of STAGES.  A
stage is one GEMM group (every MMA accumulates into one 128-column TMEM region) followed by an
EPILOGUE: a short list of micro-ops over a per-row register vector that loads the accumulator
(+ bias), optionally spills it to the skip stack, computes LayerNorm statistics, and EMITS the
next stage's A operand as fp16 (hi, lo) K-chunks into the shared-memory operand ring.

Layout decisions made here:
  * activations x live in TMEM as x = acc[region] + xb, with xb a host-precomputed cumulative
    bias vector, so the residual add `h + x` is the MMA accumulating into x's own region;
  * `cat(x, skip)` is never materialised: its K-chunks are emitted skip-part first, x-part
    second, and the weight rows are permuted to match;
  * an UpBlock's `lin3(a3) + shortcut(cat(x, skip))` is ONE GEMM group of K = D + 2D;
  * a single-token AttentionBlock is ONE accumulate stage with the fused matrix output.weight @ V.weight;
  * weights are fp16 core-matrix images ([n/8][k/8][n%8][8], one image per <=64-wide K-chunk)
    streamed by 1-D bulk TMA; with nterms == 3 a second image holds the fp16 residual of W;
  * every fp32 side parameter a stage's epilogue needs (bias, LayerNorm gamma/beta) is packed
    into ONE contiguous per-stage package that the TMA warp streams into shared memory ahead
    of the epilogue; the hoisted time bias (+ lin1.bias) comes from the per-step table row;
  * widths are padded to multiples of 16 (UMMA K / N granularity); pad rows/cols are zero.
Reference semantics: ddpm_opt/UNetCF.py:83-95, :318-356.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .unet import AttentionBlock, DownBlock, ResidualBlock, UNet1D, UpBlock

CHUNK_K = _lib.TC_VARIANT["chunk"]      # K columns per operand chunk (must match the library build)
MAX_W = _lib.TC_VARIANT["region"]       # widest vector = TMEM columns per accumulator region
PKG_MAX_FLOATS = 640

# primitive micro-ops emitted by the lowering below (an intermediate form; `_select_ops` turns each
# stage's primitive sequence into the streaming ops the kernel implements)
TE_LOAD, TE_LOAD_SKIP, TE_LOAD_INPUT, TE_STORE_SKIP, TE_STORE_OUT = 1, 2, 3, 4, 5
TE_STATS, TE_EMIT_LN, TE_EMIT_RAW, TE_EMIT_COND = 6, 7, 8, 9
STATS_RESET, STATS_FINISH = 1, 2
# streaming epilogue ops (mirror diffsg_b200/csrc/unet_tc.cuh).  Each walks its source in groups of
# 16 columns; nothing but the current group lives in registers.
OP_LN, OP_CATLN, OP_RAW_T, OP_RAW_S, OP_RAW_IN, OP_OUT = 1, 2, 3, 4, 5, 6
F_TIME, F_COND, F_PUSH, F_DEFER = 1, 2, 4, 8
NONE8 = 255

EPI_DT = np.dtype([("kind", "u1"), ("np", "u1"), ("dt", "u1"), ("misc", "u1"), ("slot", "u1"), ("off0", "u1"),
                   ("off1", "u1"), ("off2", "u1")])
CHUNK_DT = np.dtype([("kw", "u2"), ("flags", "u2"), ("w_off16", "u4")])          # w_off16: offset / 16 bytes
STAGE_DT = np.dtype([("chunk_begin", "u2"), ("epi_begin", "u2"), ("n_chunks", "u1"), ("n_epi", "u1"),
                     ("n16", "u1"), ("bits", "u1"), ("pkg_off4", "u4"), ("tt_src4", "u2"), ("pkg_f4", "u1"),
                     ("tt_f4", "u1")])
CHUNK_COND = 1
assert EPI_DT.itemsize == 8 and CHUNK_DT.itemsize == 8 and STAGE_DT.itemsize == 16


def pad16(n: int) -> int:
    return max(16, (n + 15) & ~15)


def supported(model: UNet1D) -> str | None:
    """None if the tensor-core engine can run this topology, else the reason it cannot."""
    widths = [model.proj_dim, *model.dims]
    if any(w > MAX_W or w < 1 for w in widths):     # any width: vectors are padded to multiples of 16 with exact zeros
        return f"internal widths must be in [1, {MAX_W}]"
    if model.input_dim > MAX_W or model.cond_dim > 128:
        return f"input_dim > {MAX_W} or cond_dim > 128"
    return None


@dataclass
class TcProgram:
    stages: list = field(default_factory=list)
    chunks: list = field(default_factory=list)
    epis: list = field(default_factory=list)
    wpieces: list = field(default_factory=list)     # (byte offset, n, npad, k, kw, fn -> [n, k] fp32 weight slice)
    ppieces: list = field(default_factory=list)     # (float offset, n, fn -> flat fp32)
    w_bytes: int = 0
    n_params: int = 0
    skip_widths: list = field(default_factory=list)  # padded widths
    time_blocks: list = field(default_factory=list)  # (t_off, time_emb Linear, lin1 Linear)
    tt_stride: int = 0
    input_dim: int = 0
    cond_dim: int = 0
    nterms: int = 2
    _open: dict | None = None

    # ---- per-stage parameter package
    def vec(self, fn, n: int, npad: int) -> int:
        """Append a vector to the open stage's package; returns its float4 offset inside the package."""
        st = self._open
        off = st["pkg_floats"]                      # offset inside the static (blob-resident) part
        self.ppieces.append((st["pkg_off"] + off, n, fn))
        st["pkg_floats"] += (npad + 3) & ~3
        assert st["tt_floats"] + st["pkg_floats"] <= PKG_MAX_FLOATS, st
        return (st["tt_floats"] + off) // 4         # shared-memory package = [time slice | static part]

    def time_slot(self, t_off: int, npad: int) -> int:
        """Reserve package space that the TMA fills from time_table[step][t_off : t_off + npad]."""
        st = self._open
        assert st["tt_src"] is None and st["pkg_floats"] == 0, "the time slice must be the first package entry"
        st["tt_src"], st["tt_floats"] = t_off, npad
        return 0

    def weight_chunk(self, fn, n: int, npad: int, k: int, kw: int) -> int:
        off = self.w_bytes
        self.wpieces.append((off, n, npad, k, kw, fn))
        self.w_bytes += npad * kw * 2
        return off

    # ---- stage construction
    def begin_stage(self, n_out: int, region: int, accumulate: bool, has_gemm: bool = True):
        assert self._open is None
        self._open = dict(chunk_begin=len(self.chunks), n16=pad16(n_out) // 16, region=region,
                          accumulate=int(accumulate), has_gemm=int(has_gemm), epi_begin=len(self.epis),
                          pkg_off=self.n_params, pkg_floats=0, tt_src=None, tt_floats=0)

    def add_k_segment(self, weight_fn, n_out: int, k: int, cond: bool = False):
        """Append the K-chunks of one operand segment of width k (weight_fn() -> [n_out, k])."""
        npad, kp = pad16(n_out), pad16(k)
        for k0 in range(0, kp, CHUNK_K):
            kw = min(CHUNK_K, kp - k0)
            kreal = max(0, min(k - k0, kw))
            off = self.weight_chunk(lambda k0=k0, kreal=kreal: weight_fn()[:, k0:k0 + kreal], n_out, npad, kreal, kw)
            self.chunks.append(dict(kw=kw, flags=CHUNK_COND if cond else 0, w_off16=off // 16))

    def epi(self, kind, width=16, dt=0, region=0, flags=0, slot=0, off0=NONE8, off1=NONE8, off2=NONE8):
        self.epis.append(dict(kind=kind, np=pad16(width) // 8, dt=dt, region=region, flags=flags, slot=slot,
                              off0=off0, off1=off1, off2=off2))

    def end_stage(self):
        st = self._open
        st["n_chunks"] = len(self.chunks) - st["chunk_begin"]
        st["n_epi"] = len(self.epis) - st["epi_begin"]
        # shared-memory package = [time slice (from the per-step table row) | static part (blob)]
        self.n_params += st["pkg_floats"]
        self._select_ops(st)
        self.stages.append(st)
        self._open = None

    def _select_ops(self, st):
        """Rewrite the stage's primitive op sequence into streaming ops."""
        prim = self.epis[st["epi_begin"]:]
        out, i = [], 0

        def kind(j):
            return prim[j]["kind"] if j < len(prim) else None

        def op(k, src, **kw):
            d = dict(kind=k, np=src["np"], dt=src["dt"], region=src.get("region", 0), flags=0, slot=0,
                     off0=NONE8, off1=NONE8, off2=NONE8)
            d.update(kw)
            out.append(d)

        while i < len(prim):
            o = prim[i]
            if o["kind"] == TE_LOAD_INPUT and kind(i + 1) == TE_EMIT_RAW:
                op(OP_RAW_IN, o)
                i += 2
            elif o["kind"] == TE_LOAD_SKIP and kind(i + 1) == TE_EMIT_RAW:
                op(OP_RAW_S, o, slot=o["slot"])
                i += 2
            elif o["kind"] == TE_LOAD:
                j, flags, slot = i + 1, o["flags"] & F_TIME, 0
                if kind(j) == TE_STORE_SKIP:
                    flags, slot, j = flags | F_PUSH, prim[j]["slot"], j + 1
                if kind(j) == TE_STORE_OUT:
                    op(OP_OUT, o, off0=o["off0"])
                    i = j + 1
                elif kind(j) == TE_EMIT_RAW:
                    op(OP_RAW_T, o, flags=flags, slot=slot, off0=o["off0"])
                    i = j + 1
                elif kind(j) == TE_STATS and prim[j]["flags"] == (STATS_RESET | STATS_FINISH) and kind(j + 1) == TE_EMIT_LN:
                    ln = prim[j + 1]
                    j += 2
                    if kind(j) == TE_EMIT_COND:
                        flags, j = flags | F_COND, j + 1
                    op(OP_LN, o, flags=flags, slot=slot, off0=o["off0"], off1=ln["off0"], off2=ln["off1"])
                    i = j
                elif (kind(j) == TE_STATS and kind(j + 1) == TE_LOAD_SKIP and kind(j + 2) == TE_STATS
                      and kind(j + 3) == TE_EMIT_LN and kind(j + 4) == TE_LOAD and kind(j + 5) == TE_EMIT_LN):
                    sk, ln_s, ln_x = prim[j + 1], prim[j + 3], prim[j + 5]
                    dp4 = o["np"] * 2
                    # package order fixed by _emit_next: gamma_x, beta_x, gamma_s, beta_s (contiguous)
                    assert ln_x["off1"] == ln_x["off0"] + dp4 and ln_s["off0"] == ln_x["off0"] + 2 * dp4 \
                        and ln_s["off1"] == ln_x["off0"] + 3 * dp4 and sk["np"] == o["np"]
                    assert not (flags & F_PUSH)
                    op(OP_CATLN, o, flags=flags, slot=sk["slot"], off0=o["off0"], off1=ln_x["off0"])
                    i = j + 6
                else:
                    raise AssertionError(f"no streaming op for primitive sequence at {i}: {[q['kind'] for q in prim]}")
            else:
                raise AssertionError(f"no streaming op for primitive sequence at {i}: {[q['kind'] for q in prim]}")
        self.epis[st["epi_begin"]:] = out
        st["n_epi"] = len(out)

    def mark_deferred(self):
        """An LN whose source region is overwritten by the NEXT stage's GEMM must not let that GEMM
        start before the source has been read completely: its operand chunks are published at the end."""
        for a, b in zip(self.stages, self.stages[1:]):
            if not b["has_gemm"]:
                continue
            for o in self.epis[a["epi_begin"]:a["epi_begin"] + a["n_epi"]]:
                if o["kind"] in (OP_LN, OP_RAW_T) and o["region"] == b["region"]:
                    o["flags"] |= F_DEFER

    # ---- arrays for the C-ABI
    def arrays(self):
        s = np.zeros(len(self.stages), STAGE_DT)
        for i, d in enumerate(self.stages):
            s[i] = (d["chunk_begin"], d["epi_begin"], d["n_chunks"], d["n_epi"], d["n16"],
                    d["region"] | (d["accumulate"] << 1) | (d["has_gemm"] << 2) | ((d["tt_src"] is not None) << 3),
                    d["pkg_off"] // 4, (d["tt_src"] or 0) // 4, d["pkg_floats"] // 4, d["tt_floats"] // 4)
        c = np.zeros(len(self.chunks), CHUNK_DT)
        for i, d in enumerate(self.chunks):
            c[i] = (d["kw"], d["flags"], d["w_off16"])
        e = np.zeros(len(self.epis), EPI_DT)
        for i, d in enumerate(self.epis):
            e[i] = (d["kind"], d["np"], d["dt"], d["region"] | (d["flags"] << 1), d["slot"], d["off0"], d["off1"], d["off2"])
        return s, c, e

    def gemm_macs(self):
        """(x-path, cond) algorithmic MACs per row-forward from the un-padded shapes."""
        x = c = 0
        for (_, n, _, k, _, _), ch in zip(self.wpieces, self.chunks):
            if ch["flags"] & CHUNK_COND:
                c += n * k
            else:
                x += n * k
        return x, c


def _emit_next(p: TcProgram, plan, xr, xb_off4, width):
    """Epilogue tail shared by every stage that ends with a finished activation x (in region xr,
    cumulative bias at package offset xb_off4): what it must EMIT depends on the module that consumes x next."""
    dp = pad16(width)
    kind = plan[0]
    if kind in ("res", "final"):   # identity-shortcut ResidualBlock / output head: LN -> Swish
        norm = plan[1].norm1 if kind == "res" else plan[1]
        p.epi(TE_STATS, width=dp, dt=width, flags=STATS_RESET | STATS_FINISH)
        g = p.vec(lambda: norm.weight, width, dp)
        b = p.vec(lambda: norm.bias, width, dp)
        p.epi(TE_EMIT_LN, width=dp, dt=width, off0=g, off1=b)
    elif kind == "raw":            # Down/Upsample Linear consumes x itself
        p.epi(TE_EMIT_RAW, width=dp, dt=width)
    elif kind == "up":             # UpBlock: LN1 over cat(x, skip); chunks: skip part, then x part
        blk, slot = plan[1], plan[2]
        g_x = p.vec(lambda: blk.norm1.weight[:width], width, dp)
        b_x = p.vec(lambda: blk.norm1.bias[:width], width, dp)
        g_s = p.vec(lambda: blk.norm1.weight[width:], width, dp)
        b_s = p.vec(lambda: blk.norm1.bias[width:], width, dp)
        p.epi(TE_STATS, width=dp, dt=width, flags=STATS_RESET)
        p.epi(TE_LOAD_SKIP, width=dp, dt=width, slot=slot)
        p.epi(TE_STATS, width=dp, dt=width, flags=STATS_FINISH)
        p.epi(TE_EMIT_LN, width=dp, dt=width, off0=g_s, off1=b_s)
        p.epi(TE_LOAD, width=dp, dt=width, region=xr, off0=xb_off4)
        p.epi(TE_EMIT_LN, width=dp, dt=width, off0=g_x, off1=b_x)
    else:
        raise ValueError(kind)


def lower_tc(model: UNet1D, nterms: int = 2) -> TcProgram:
    why = supported(model)
    if why:
        raise ValueError(f"tensor-core engine does not support this UNet1D: {why}")
    p = TcProgram(input_dim=model.input_dim, cond_dim=model.cond_dim, nterms=nterms)
    M, C = model.input_dim, model.cond_dim

    # flat module sequence with look-ahead ("what consumes x next?"); `push`: the module's output is a skip
    seq, pushes = [], []

    def add(kind, mod, push=False):
        seq.append((kind, mod))
        pushes.append(push)

    def add_res(kind, stage_mod, res, push):
        att = getattr(stage_mod, "attn", None)
        if isinstance(att, AttentionBlock):     # single-token attention = x + output(V(x)): one more accumulate stage
            add(kind, res)
            add("attn", att, push)
        else:
            add(kind, res, push)

    for m in model.down:
        if isinstance(m, DownBlock):
            add_res("down_res", m, m.res, True)
        else:
            add("lin", m.lin, True)
    add_res("mid_res", model.middle, model.middle.res1, False)
    add("mid_res", model.middle.res2)
    for m in model.up:
        if isinstance(m, UpBlock):
            add_res("up_res", m, m.res, False)
        else:
            add("lin", m.lin)
    add("final", None)

    # skip slots are pushed after feature_proj and after every `down` module, popped LIFO by UpBlocks
    n_push = 1 + len(model.down)
    assert sum(pushes) == len(model.down)
    pops = [i for i, (k, _) in enumerate(seq) if k == "up_res"]
    assert len(pops) == n_push
    pop_slot = {idx: n_push - 1 - j for j, idx in enumerate(pops)}

    def consumer_plan(i):
        k, mod = seq[i]
        if k in ("down_res", "mid_res"):
            return ("res", mod)
        if k in ("lin", "attn"):
            return ("raw",)
        if k == "up_res":
            return ("up", mod, pop_slot[i])
        return ("final", model.norm)

    def sum_fn(fns):
        fns = tuple(fns)
        return lambda: sum(f() for f in fns)

    # ---- stage 0: operand of feature_proj = the raw input row
    p.begin_stage(16, 0, False, has_gemm=False)
    p.epi(TE_LOAD_INPUT, width=pad16(M), dt=M)
    p.epi(TE_EMIT_RAW, width=pad16(M), dt=M)
    p.end_stage()

    # ---- feature_proj
    fp = model.feature_proj
    width = model.proj_dim
    xr = 0
    p.begin_stage(width, xr, False)
    p.add_k_segment(lambda: fp.weight, width, M)
    xb = [lambda: fp.bias]                 # bias terms summed into the cumulative vector of x
    xoff = p.vec(sum_fn(xb), width, pad16(width))
    p.epi(TE_LOAD, width=width, dt=width, region=xr, off0=xoff)
    slot = 0
    p.skip_widths.append(pad16(width))
    p.epi(TE_STORE_SKIP, width=width, dt=width, slot=slot)
    _emit_next(p, consumer_plan(0), xr, xoff, width)
    p.end_stage()
    slot += 1

    for i, (kind, mod) in enumerate(seq):
        nxt = consumer_plan(i + 1) if i + 1 < len(seq) else None
        if kind == "lin":
            lin = mod
            hr = 1 - xr
            p.begin_stage(lin.out_features, hr, False)
            p.add_k_segment(lambda lin=lin: lin.weight, lin.out_features, lin.in_features)
            width = lin.out_features
            xr = hr
            xb = [lambda lin=lin: lin.bias]
            xoff = p.vec(sum_fn(xb), width, pad16(width))
            p.epi(TE_LOAD, width=width, dt=width, region=xr, off0=xoff)
            if pushes[i]:
                p.skip_widths.append(pad16(width))
                p.epi(TE_STORE_SKIP, width=width, dt=width, slot=slot)
                slot += 1
            _emit_next(p, nxt, xr, xoff, width)
            p.end_stage()
        elif kind == "attn":
            # x' = x + output(V(x)) (the reference's length-1 sequence makes softmax == 1, UNetCF.py:123-157):
            # ONE accumulate stage with the host-fused matrix output.weight . V.weight, bias folded into xb
            att = mod
            dk, nh = att.d_k, att.n_heads

            def v_rows(att=att, dk=dk, nh=nh):
                return torch.cat([torch.arange(h * 3 * dk + 2 * dk, (h + 1) * 3 * dk) for h in range(nh)])

            p.begin_stage(width, xr, True)
            p.add_k_segment(lambda att=att, vr=v_rows: att.output.weight.double().matmul(att.projection.weight[vr()].double()).float(),
                            width, width)
            xb = xb + [lambda att=att, vr=v_rows: (att.output.weight.double().matmul(att.projection.bias[vr()].double())
                                                   + att.output.bias.double()).float()]
            xoff = p.vec(sum_fn(xb), width, pad16(width))
            p.epi(TE_LOAD, width=width, dt=width, region=xr, off0=xoff)
            if pushes[i]:
                p.skip_widths.append(pad16(width))
                p.epi(TE_STORE_SKIP, width=width, dt=width, slot=slot)
                slot += 1
            _emit_next(p, nxt, xr, xoff, width)
            p.end_stage()
        elif kind in ("down_res", "mid_res", "up_res"):
            blk: ResidualBlock = mod
            dout = blk.out_dim
            dp = pad16(dout)
            hr = 1 - xr
            is_up = kind == "up_res"
            t_off = p.tt_stride
            p.time_blocks.append((t_off, blk.time_emb, blk.lin1))
            p.tt_stride += dp
            # ---- G1: h = lin1(a1) + [b1 + time]  (bias = the per-step table slice)
            p.begin_stage(dout, hr, False)
            if is_up:   # K order: skip part (cat columns [width:]) then x part (cat columns [:width])
                p.add_k_segment(lambda blk=blk, w=width: blk.lin1.weight[:, w:], dout, blk.in_dim - width)
                p.add_k_segment(lambda blk=blk, w=width: blk.lin1.weight[:, :w], dout, width)
            else:
                p.add_k_segment(lambda blk=blk: blk.lin1.weight, dout, blk.in_dim)
            p.epi(TE_LOAD, width=dout, dt=dout, region=hr, flags=F_TIME, off0=p.time_slot(t_off, dp))
            p.epi(TE_STATS, width=dout, dt=dout, flags=STATS_RESET | STATS_FINISH)
            g = p.vec(lambda blk=blk: blk.norm2.weight, dout, dp)
            b = p.vec(lambda blk=blk: blk.norm2.bias, dout, dp)
            p.epi(TE_EMIT_LN, width=dout, dt=dout, off0=g, off1=b)
            p.epi(TE_EMIT_COND)
            p.end_stage()
            # ---- G2: h = lin2(a2) + b2 + cond_emb(swish(cond))
            p.begin_stage(dout, hr, False)
            p.add_k_segment(lambda blk=blk: blk.lin2.weight, dout, dout)
            p.add_k_segment(lambda blk=blk: blk.cond_emb.weight, dout, C, cond=True)
            b2 = p.vec(lambda blk=blk: blk.lin2.bias + blk.cond_emb.bias, dout, dp)
            p.epi(TE_LOAD, width=dout, dt=dout, region=hr, off0=b2)
            p.epi(TE_STATS, width=dout, dt=dout, flags=STATS_RESET | STATS_FINISH)
            g = p.vec(lambda blk=blk: blk.norm3.weight, dout, dp)
            b = p.vec(lambda blk=blk: blk.norm3.bias, dout, dp)
            p.epi(TE_EMIT_LN, width=dout, dt=dout, off0=g, off1=b)
            if is_up:   # raw operands of the shortcut Linear: x part, then skip part
                sw = blk.in_dim - width
                p.epi(TE_LOAD, width=width, dt=width, region=xr, off0=p.vec(sum_fn(xb), width, pad16(width)))
                p.epi(TE_EMIT_RAW, width=width, dt=width)
                p.epi(TE_LOAD_SKIP, width=sw, dt=sw, slot=pop_slot[i])
                p.epi(TE_EMIT_RAW, width=sw, dt=sw)
            p.end_stage()
            # ---- G3: x' = lin3(a3) + b3 + shortcut(x)
            if is_up:
                p.begin_stage(dout, hr, False)
                p.add_k_segment(lambda blk=blk: blk.lin3.weight, dout, dout)
                p.add_k_segment(lambda blk=blk, w=width: blk.shortcut.weight[:, :w], dout, width)
                p.add_k_segment(lambda blk=blk, w=width: blk.shortcut.weight[:, w:], dout, blk.in_dim - width)
                xr = hr
                xb = [lambda blk=blk: blk.lin3.bias, lambda blk=blk: blk.shortcut.bias]
            else:
                assert isinstance(blk.shortcut, nn.Identity) and blk.in_dim == dout
                p.begin_stage(dout, xr, True)
                p.add_k_segment(lambda blk=blk: blk.lin3.weight, dout, dout)
                xb = xb + [lambda blk=blk: blk.lin3.bias]
            width = dout
            xoff = p.vec(sum_fn(xb), width, dp)
            p.epi(TE_LOAD, width=width, dt=width, region=xr, off0=xoff)
            if pushes[i]:
                p.skip_widths.append(dp)
                p.epi(TE_STORE_SKIP, width=width, dt=width, slot=slot)
                slot += 1
            _emit_next(p, nxt, xr, xoff, width)
            p.end_stage()
        elif kind == "final":
            fin = model.final
            hr = 1 - xr
            p.begin_stage(M, hr, False)
            p.add_k_segment(lambda: fin.weight, M, fin.in_features)
            p.epi(TE_LOAD, width=M, dt=M, region=hr, off0=p.vec(lambda: fin.bias, M, pad16(M)))
            p.epi(TE_STORE_OUT, width=M, dt=M)
            p.end_stage()
    assert slot == n_push, (slot, n_push)
    p.mark_deferred()
    return p


def pack_tc_weights(p: TcProgram, device):
    """-> (w_hi fp16 blob, w_lo fp16 blob or None, params fp32 blob) on `device`."""
    with torch.no_grad():
        hi = torch.zeros(max(p.w_bytes // 2, 8), dtype=torch.float16, device=device)
        lo = torch.zeros_like(hi) if p.nterms >= 3 else None
        for off, n, npad, k, kw, fn in p.wpieces:
            w = torch.zeros(npad, kw, dtype=torch.float32, device=device)
            if k > 0:
                w[:n, :k] = fn().detach().to(device=device, dtype=torch.float32)
            wh = w.to(torch.float16)
            img = lambda t: t.reshape(npad // 8, 8, kw // 8, 8).permute(0, 2, 1, 3).reshape(-1)
            hi[off // 2:off // 2 + npad * kw] = img(wh)
            if lo is not None:
                lo[off // 2:off // 2 + npad * kw] = img((w - wh.to(torch.float32)).to(torch.float16))
        params = torch.zeros(max(p.n_params, 4), dtype=torch.float32, device=device)
        for off, n, fn in p.ppieces:
            params[off:off + n] = fn().detach().to(device=device, dtype=torch.float32).reshape(-1)
    return hi, lo, params


def time_table_tc(model: UNet1D, p: TcProgram, t_values: torch.Tensor) -> torch.Tensor:
    """Hoisted time path for this program's t_off layout, with each block's lin1.bias folded in:
    row r, block k -> lin1_k.bias + time_emb_k(Swish(TimeEmbedding(t_values[r])))."""
    from .packer import _swish, sinusoid
    te = model.time_emb
    F = torch.nn.functional
    with torch.no_grad():
        e = sinusoid(t_values, model.proj_dim)
        e = F.linear(_swish(F.linear(e, te.lin1.weight, te.lin1.bias)), te.lin2.weight, te.lin2.bias)
        a = _swish(e)
        tab = torch.zeros(a.shape[0], max(p.tt_stride, 4), dtype=torch.float32, device=a.device)
        for t_off, lin, lin1 in p.time_blocks:
            tab[:, t_off:t_off + lin.out_features] = F.linear(a, lin.weight, lin.bias) + lin1.bias
    return tab.contiguous()
