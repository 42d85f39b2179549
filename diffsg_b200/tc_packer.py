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
  * biases ride on the tensor cores: every GEMM group ends with one K = 16 "bias MMA" whose A operand is the
    kernel's constant tile of ones (K columns 0..2) and whose W image holds the bias as three fp16 terms
    (hi, mid, lo; exact to 2^-33); the image travels with the group's last weight chunk (same W-ring stage).
    Static biases live in the weight blob; the hoisted time bias (lin1.bias + time_emb(Swish(TimeEmbedding(t))))
    is one image row per reverse step (`time_images`).  `UNet1D.forward` (rows with arbitrary t) skips the
    time-bias MMAs and adds the row's fp32 table slice in the epilogue instead;
  * the LayerNorm gamma / beta a stage's epilogue needs are packed into ONE contiguous per-stage package that
    the TMA warp streams into shared memory ahead of the epilogue;
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
PKG_MAX_FLOATS = 512
BIAS_K = 16                             # K of a bias chunk; K columns 0..2 carry the three fp16 terms

# streaming epilogue ops (mirror diffsg_b200/csrc/unet_tc.cuh).  Each walks its source in groups of
# 16 columns; nothing but the current group lives in registers.
OP_LN, OP_CATLN, OP_RAW_T, OP_RAW_S, OP_RAW_IN, OP_OUT = 1, 2, 3, 4, 5, 6
F_TIME, F_COND, F_PUSH, F_DEFER = 1, 2, 4, 8

EPI_DT = np.dtype([("kind", "u1"), ("np", "u1"), ("dt", "u1"), ("misc", "u1"), ("slot", "u1"), ("off1", "u1"),
                   ("tt_src4", "u2")])
CHUNK_DT = np.dtype([("kw", "u2"), ("flags", "u2"), ("w_off16", "u4")])          # w_off16: offset / 16 bytes
STAGE_DT = np.dtype([("chunk_begin", "u2"), ("epi_begin", "u2"), ("n_chunks", "u1"), ("n_epi", "u1"),
                     ("n16", "u1"), ("bits", "u1"), ("pkg_off16", "u2"), ("pkg_f4", "u1"), ("pad_", "u1"),
                     ("bias_off16", "u4")])      # bits: region | accumulate << 1 | has_gemm << 2 | time_bias << 3
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
    bpieces: list = field(default_factory=list)     # (byte offset, n, npad, fn -> [n] fp32 bias): static bias chunks
    ppieces: list = field(default_factory=list)     # (float offset, n, fn -> flat fp32): LayerNorm gamma / beta
    w_bytes: int = 0
    n_params: int = 0
    skip_widths: list = field(default_factory=list)  # padded widths
    time_blocks: list = field(default_factory=list)  # (t_off floats, img_off bytes, npad, time_emb Linear, lin1 Linear)
    tt_stride: int = 0                               # floats per row of the fp32 time table
    img_stride: int = 0                              # bytes per row of the step image table
    input_dim: int = 0
    cond_dim: int = 0
    nterms: int = 2
    _open: dict | None = None

    # ---- per-stage parameter package (LayerNorm gamma / beta)
    def vec(self, fn, n: int, npad: int) -> int:
        """Append a vector to the open stage's package; returns its float4 offset inside the package."""
        st = self._open
        off = st["pkg_floats"]
        self.ppieces.append((st["pkg_off"] + off, n, fn))
        st["pkg_floats"] += npad
        assert npad % 16 == 0 and st["pkg_floats"] <= PKG_MAX_FLOATS, st
        return off // 4

    def ln_params(self, norm_w, norm_b, width: int) -> int:
        """gamma | beta (each padded with zeros: pad columns then come out of LayerNorm -> Swish as exact zeros)."""
        dp = pad16(width)
        off = self.vec(norm_w, width, dp)
        self.vec(norm_b, width, dp)
        return off

    # ---- stage construction
    def begin_stage(self, n_out: int, region: int, accumulate: bool, has_gemm: bool = True):
        assert self._open is None
        self._open = dict(chunk_begin=len(self.chunks), n16=pad16(n_out) // 16, region=region, n_out=n_out,
                          accumulate=int(accumulate), has_gemm=int(has_gemm), epi_begin=len(self.epis),
                          pkg_off=self.n_params, pkg_floats=0, has_bias=False, time_bias=0, bias_off16=0)

    def add_k_segment(self, weight_fn, n_out: int, k: int, cond: bool = False):
        """Append the K-chunks of one operand segment of width k (weight_fn() -> [n_out, k])."""
        npad, kp = pad16(n_out), pad16(k)
        for k0 in range(0, kp, CHUNK_K):
            kw = min(CHUNK_K, kp - k0)
            kreal = max(0, min(k - k0, kw))
            off = self.w_bytes
            self.wpieces.append((off, n_out, npad, kreal, kw, lambda k0=k0, kreal=kreal: weight_fn()[:, k0:k0 + kreal]))
            self.w_bytes += npad * kw * 2
            self.chunks.append(dict(kw=kw, flags=CHUNK_COND if cond else 0, w_off16=off // 16, macs=n_out * kreal))

    def add_bias(self, bias_fn):
        """The GEMM group's static bias (image in the weight blob)."""
        st = self._open
        assert not st["has_bias"]
        npad = st["n16"] * 16
        off = self.w_bytes
        self.bpieces.append((off, st["n_out"], npad, bias_fn))
        self.w_bytes += npad * BIAS_K * 2
        st.update(has_bias=True, time_bias=0, bias_off16=off // 16)

    def add_time_bias(self, time_emb, lin1) -> int:
        """The GEMM group's per-step bias lin1.bias + time_emb(.): one image per reverse step in the step image
        table; returns the column offset of the same values in the fp32 time table."""
        st = self._open
        assert not st["has_bias"]
        npad = st["n16"] * 16
        t_off, img_off = self.tt_stride, self.img_stride
        self.time_blocks.append((t_off, img_off, npad, time_emb, lin1))
        self.tt_stride += npad
        self.img_stride += npad * BIAS_K * 2
        st.update(has_bias=True, time_bias=1, bias_off16=img_off // 16)
        return t_off

    def epi(self, kind, width, region=0, flags=0, slot=0, off1=0, tt_src=0):
        assert 1 <= width <= MAX_W and tt_src % 4 == 0
        self.epis.append(dict(kind=kind, np=pad16(width) // 8, dt=width, region=region, flags=flags, slot=slot,
                              off1=off1, tt_src4=tt_src // 4))

    def end_stage(self):
        st = self._open
        assert not st["has_gemm"] or st["has_bias"], "every GEMM group carries a bias image"
        assert st["pkg_off"] % 16 == 0 and st["pkg_floats"] // 4 < 256
        st["n_chunks"] = len(self.chunks) - st["chunk_begin"]
        st["n_epi"] = len(self.epis) - st["epi_begin"]
        self.n_params += st["pkg_floats"]
        self.stages.append(st)
        self._open = None

    def mark_deferred(self):
        """An op whose source region is overwritten by the NEXT stage's GEMM must not let that GEMM
        start before the source has been read completely: its operand chunks are published at the end."""
        for a, b in zip(self.stages, self.stages[1:]):
            if not b["has_gemm"]:
                continue
            for o in self.epis[a["epi_begin"]:a["epi_begin"] + a["n_epi"]]:
                if o["kind"] in (OP_LN, OP_RAW_T) and o["region"] == b["region"]:
                    o["flags"] |= F_DEFER
                assert not (o["kind"] == OP_CATLN and o["region"] == b["region"]), "cat-LN source overwritten by the next GEMM"

    # ---- arrays for the C-ABI
    def arrays(self):
        s = np.zeros(len(self.stages), STAGE_DT)
        for i, d in enumerate(self.stages):
            s[i] = (d["chunk_begin"], d["epi_begin"], d["n_chunks"], d["n_epi"], d["n16"],
                    d["region"] | (d["accumulate"] << 1) | (d["has_gemm"] << 2) | (d["time_bias"] << 3),
                    d["pkg_off"] // 16, d["pkg_floats"] // 4, 0, d["bias_off16"])
        c = np.zeros(len(self.chunks), CHUNK_DT)
        for i, d in enumerate(self.chunks):
            c[i] = (d["kw"], d["flags"], d["w_off16"])
        e = np.zeros(len(self.epis), EPI_DT)
        for i, d in enumerate(self.epis):
            e[i] = (d["kind"], d["np"], d["dt"], d["region"] | (d["flags"] << 1), d["slot"], d["off1"], d["tt_src4"])
        return s, c, e

    def gemm_macs(self):
        """(x-path, cond) algorithmic MACs per row-forward from the un-padded shapes (biases are not MACs)."""
        x = sum(ch["macs"] for ch in self.chunks if not ch["flags"] & CHUNK_COND)
        c = sum(ch["macs"] for ch in self.chunks if ch["flags"] & CHUNK_COND)
        return x, c


def _emit_next(p: TcProgram, plan, xr, width, push_slot=None):
    """Epilogue of every stage that ends with a finished activation x (in TMEM region xr, bias included): what it
    must EMIT depends on the module that consumes x next; `push_slot`: x is also a skip (spilled with its moments)."""
    kind = plan[0]
    flags = F_PUSH if push_slot is not None else 0
    slot = push_slot if push_slot is not None else 0
    if kind in ("res", "final"):   # identity-shortcut ResidualBlock / output head: LN -> Swish
        norm = plan[1].norm1 if kind == "res" else plan[1]
        off = p.ln_params(lambda: norm.weight, lambda: norm.bias, width)
        p.epi(OP_LN, width, region=xr, flags=flags, slot=slot, off1=off)
    elif kind == "raw":            # Down/Upsample Linear (or the fused attention matrix) consumes x itself
        p.epi(OP_RAW_T, width, region=xr, flags=flags, slot=slot)
    elif kind == "up":             # UpBlock: LN1 over cat(x, skip); chunks: skip part, then x part
        blk, pop = plan[1], plan[2]
        assert push_slot is None and blk.in_dim == 2 * width and p.skip_widths[pop] == pad16(width), \
            "cat LayerNorm expects a skip as wide as x"
        off = p.ln_params(lambda: blk.norm1.weight[:width], lambda: blk.norm1.bias[:width], width)
        p.ln_params(lambda: blk.norm1.weight[width:], lambda: blk.norm1.bias[width:], width)
        p.epi(OP_CATLN, width, region=xr, slot=pop, off1=off)
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

    slot = 0

    def finish(i_next, xr, width, push):
        """Close the stage that produced x: push it if it is a skip, emit what its consumer needs."""
        nonlocal slot
        push_slot = None
        if push:
            p.skip_widths.append(pad16(width))
            push_slot = slot
            slot += 1
        _emit_next(p, consumer_plan(i_next), xr, width, push_slot)
        p.end_stage()

    # ---- stage 0: operand of feature_proj = the raw input row
    p.begin_stage(16, 0, False, has_gemm=False)
    p.epi(OP_RAW_IN, M)
    p.end_stage()

    # ---- feature_proj
    fp = model.feature_proj
    width = model.proj_dim
    xr = 0
    p.begin_stage(width, xr, False)
    p.add_k_segment(lambda: fp.weight, width, M)
    p.add_bias(lambda: fp.bias)
    finish(0, xr, width, True)

    for i, (kind, mod) in enumerate(seq):
        if kind == "lin":
            lin = mod
            xr = 1 - xr
            p.begin_stage(lin.out_features, xr, False)
            p.add_k_segment(lambda lin=lin: lin.weight, lin.out_features, lin.in_features)
            p.add_bias(lambda lin=lin: lin.bias)
            width = lin.out_features
            finish(i + 1, xr, width, pushes[i])
        elif kind == "attn":
            # x' = x + output(V(x)) (the reference's length-1 sequence makes softmax == 1, UNetCF.py:123-157):
            # ONE accumulate stage with the host-fused matrix output.weight . V.weight and the fused bias
            att = mod
            dk, nh = att.d_k, att.n_heads

            def v_rows(att=att, dk=dk, nh=nh):
                return torch.cat([torch.arange(h * 3 * dk + 2 * dk, (h + 1) * 3 * dk) for h in range(nh)])

            p.begin_stage(width, xr, True)
            p.add_k_segment(lambda att=att, vr=v_rows: att.output.weight.double().matmul(att.projection.weight[vr()].double()).float(),
                            width, width)
            p.add_bias(lambda att=att, vr=v_rows: (att.output.weight.double().matmul(att.projection.bias[vr()].double())
                                                   + att.output.bias.double()).float())
            finish(i + 1, xr, width, pushes[i])
        elif kind in ("down_res", "mid_res", "up_res"):
            blk: ResidualBlock = mod
            dout = blk.out_dim
            hr = 1 - xr
            is_up = kind == "up_res"
            # ---- G1: h = lin1(a1) + [b1 + time]  (bias chunk = this step's image of the time table)
            p.begin_stage(dout, hr, False)
            if is_up:   # K order: skip part (cat columns [width:]) then x part (cat columns [:width])
                p.add_k_segment(lambda blk=blk, w=width: blk.lin1.weight[:, w:], dout, blk.in_dim - width)
                p.add_k_segment(lambda blk=blk, w=width: blk.lin1.weight[:, :w], dout, width)
            else:
                p.add_k_segment(lambda blk=blk: blk.lin1.weight, dout, blk.in_dim)
            t_off = p.add_time_bias(blk.time_emb, blk.lin1)
            off = p.ln_params(lambda blk=blk: blk.norm2.weight, lambda blk=blk: blk.norm2.bias, dout)
            p.epi(OP_LN, dout, region=hr, flags=F_TIME | F_COND, off1=off, tt_src=t_off)
            p.end_stage()
            # ---- G2: h = lin2(a2) + b2 + cond_emb(swish(cond))   (uncond pass: Swish(0) = 0 -> the bias alone)
            p.begin_stage(dout, hr, False)
            p.add_k_segment(lambda blk=blk: blk.lin2.weight, dout, dout)
            p.add_k_segment(lambda blk=blk: blk.cond_emb.weight, dout, C, cond=True)
            p.add_bias(lambda blk=blk: blk.lin2.bias + blk.cond_emb.bias)
            off = p.ln_params(lambda blk=blk: blk.norm3.weight, lambda blk=blk: blk.norm3.bias, dout)
            p.epi(OP_LN, dout, region=hr, off1=off)
            if is_up:   # raw operands of the shortcut Linear: x part, then skip part
                p.epi(OP_RAW_T, width, region=xr)
                p.epi(OP_RAW_S, blk.in_dim - width, slot=pop_slot[i])
            p.end_stage()
            # ---- G3: x' = lin3(a3) + b3 + shortcut(x)
            if is_up:
                p.begin_stage(dout, hr, False)
                p.add_k_segment(lambda blk=blk: blk.lin3.weight, dout, dout)
                p.add_k_segment(lambda blk=blk, w=width: blk.shortcut.weight[:, :w], dout, width)
                p.add_k_segment(lambda blk=blk, w=width: blk.shortcut.weight[:, w:], dout, blk.in_dim - width)
                p.add_bias(lambda blk=blk: blk.lin3.bias + blk.shortcut.bias)
                xr = hr
            else:
                assert isinstance(blk.shortcut, nn.Identity) and blk.in_dim == dout
                p.begin_stage(dout, xr, True)      # the residual add is the MMA accumulating into x's own region
                p.add_k_segment(lambda blk=blk: blk.lin3.weight, dout, dout)
                p.add_bias(lambda blk=blk: blk.lin3.bias)
            width = dout
            finish(i + 1, xr, width, pushes[i])
        elif kind == "final":
            fin = model.final
            hr = 1 - xr
            p.begin_stage(M, hr, False)
            p.add_k_segment(lambda: fin.weight, M, fin.in_features)
            p.add_bias(lambda: fin.bias)
            p.epi(OP_OUT, M, region=hr)
            p.end_stage()
    assert slot == n_push, (slot, n_push)
    p.mark_deferred()
    return p


def _w_image(t: torch.Tensor, npad: int, kw: int) -> torch.Tensor:
    """[npad, kw] -> UMMA K-major no-swizzle core-matrix order [n/8][k/8][n%8][8], flat."""
    return t.reshape(npad // 8, 8, kw // 8, 8).permute(0, 2, 1, 3).reshape(-1)


def split3(b: torch.Tensor):
    """fp32 -> three fp16 terms with b == h0 + h1 + h2 up to 2^-33 |b| (and >= 2^-25 absolute: fp16 subnormals)."""
    h0 = b.to(torch.float16)
    r1 = b - h0.to(torch.float32)
    h1 = r1.to(torch.float16)
    h2 = (r1 - h1.to(torch.float32)).to(torch.float16)
    return h0, h1, h2


def bias_image(b: torch.Tensor, npad: int) -> torch.Tensor:
    """Bias vectors [..., n] -> bias-chunk images [..., npad * 16] fp16: W rows of a K = 16 chunk whose K columns
    0..2 hold the three fp16 terms (the kernel's constant A tile has ones exactly there)."""
    lead = b.shape[:-1]
    v = torch.zeros(*lead, npad, dtype=torch.float32, device=b.device)
    v[..., :b.shape[-1]] = b.to(torch.float32)
    img = torch.zeros(*lead, npad // 8, BIAS_K // 8, 8, 8, dtype=torch.float16, device=b.device)
    for j, h in enumerate(split3(v)):
        img[..., 0, :, j] = h.reshape(*lead, npad // 8, 8)
    return img.reshape(*lead, npad * BIAS_K)


def pack_tc_weights(p: TcProgram, device):
    """-> (w_hi fp16 blob incl. the static bias images, w_lo fp16 blob or None, LayerNorm params fp32 blob) on `device`."""
    with torch.no_grad():
        hi = torch.zeros(max(p.w_bytes // 2, 8), dtype=torch.float16, device=device)
        lo = torch.zeros_like(hi) if p.nterms >= 3 else None
        for off, n, npad, k, kw, fn in p.wpieces:
            w = torch.zeros(npad, kw, dtype=torch.float32, device=device)
            if k > 0:
                w[:n, :k] = fn().detach().to(device=device, dtype=torch.float32)
            wh = w.to(torch.float16)
            hi[off // 2:off // 2 + npad * kw] = _w_image(wh, npad, kw)
            if lo is not None:
                lo[off // 2:off // 2 + npad * kw] = _w_image((w - wh.to(torch.float32)).to(torch.float16), npad, kw)
        for off, n, npad, fn in p.bpieces:
            hi[off // 2:off // 2 + npad * BIAS_K] = bias_image(fn().detach().to(device=device, dtype=torch.float32).reshape(-1), npad)
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
        for t_off, _, _, lin, lin1 in p.time_blocks:
            tab[:, t_off:t_off + lin.out_features] = F.linear(a, lin.weight, lin.bias) + lin1.bias
    return tab.contiguous()


def time_images(p: TcProgram, table: torch.Tensor) -> torch.Tensor:
    """fp32 time table [R, tt_stride] -> step image table [R, img_stride / 2] fp16: per block the bias-chunk image
    of that row's slice (what the sampler's TMA warp streams as the last chunk of every lin1 GEMM group)."""
    with torch.no_grad():
        img = torch.zeros(table.shape[0], max(p.img_stride // 2, 8), dtype=torch.float16, device=table.device)
        for t_off, img_off, npad, _, _ in p.time_blocks:
            img[:, img_off // 2:img_off // 2 + npad * BIAS_K] = bias_image(table[:, t_off:t_off + npad], npad)
    return img.contiguous()
