"""Test-only CPU interpreter of the tensor-core stage program (diffsg_b200.tc_packer): same
dataflow as diffsg_b200/csrc/unet_tc.cuh (TMEM regions, operand chunk queue, bias chunks on a
constant ones tile, per-stage LayerNorm packages, skip moments stored at push time), evaluated in
fp32 (optionally with the fp16 operand rounding of the real engine)."""
import torch

from diffsg_b200 import tc_packer as T


def run_tc_program(p, w_hi, w_lo, params, table, x, t_idx, cond, mask, emulate_fp16=False, images=None):
    """`images`: step image table (sampler mode: the time-bias chunks are issued, row = t_idx which must then be
    one value); None: forward mode (time chunks skipped, the fp32 table slice of each row is added in the epilogue)."""
    B = x.shape[0]
    regions = [torch.zeros(B, 128), torch.zeros(B, 128)]
    skips, skip_stats = {}, {}
    c = cond * mask
    cpad = torch.zeros(B, T.pad16(p.cond_dim))
    cpad[:, :p.cond_dim] = c * torch.sigmoid(c)
    queue = []                       # emitted A chunks, FIFO
    out = None

    def split(a):
        if not emulate_fp16:
            return a
        hi = a.half().float()
        return hi + (a - hi).half().float()

    def emit(vec, dp):
        for k0 in range(0, dp, T.CHUNK_K):
            queue.append(split(vec[:, k0:min(dp, k0 + T.CHUNK_K)].clone()))

    def moments(v, dt):
        """shifted one-pass moments, as the kernel accumulates them"""
        shift = v[:, 0]
        d = v[:, :dt] - shift[:, None]
        return d.sum(dim=1), (d * d).sum(dim=1), shift

    def ln_emit(v, dp, mean, rstd, g, b):
        t = (v[:, :dp] * rstd[:, None] - (mean * rstd)[:, None]) * g + b
        emit(t * torch.sigmoid(t), dp)          # pad columns: gamma = beta = 0 -> exact zeros

    def bias_from_image(img, N):
        """[.., N * 16] image rows -> [.., N] bias = sum of the three fp16 terms in K columns 0..2"""
        w = img.float().reshape(*img.shape[:-1], N // 8, T.BIAS_K // 8, 8, 8)
        return w[..., 0, :, :3].sum(dim=-1).reshape(*img.shape[:-1], N)

    for st in p.stages:
        if st["has_gemm"]:
            N = st["n16"] * 16
            acc = regions[st["region"]][:, :N].clone() if st["accumulate"] else torch.zeros(B, N)
            chs = p.chunks[st["chunk_begin"]:st["chunk_begin"] + st["n_chunks"]]
            boff = st["bias_off16"] * 8
            if not st["time_bias"]:
                acc = acc + bias_from_image(w_hi[boff:boff + N * T.BIAS_K], N)[None, :]
            elif images is not None:        # forward mode (images None): the epilogue adds the row's table slice
                acc = acc + bias_from_image(images[t_idx][:, boff:boff + N * T.BIAS_K], N)
            for ch in chs:
                kw = ch["kw"]
                off = ch["w_off16"] * 8
                a = queue.pop(0)
                assert a.shape[1] == kw, (a.shape, kw)
                img = w_hi[off:off + N * kw].float()
                if p.nterms >= 3:
                    img = img + w_lo[off:off + N * kw].float()
                W = img.reshape(N // 8, kw // 8, 8, 8).permute(0, 2, 1, 3).reshape(N, kw)
                acc = acc + a @ W.t()
            regions[st["region"]][:, :N] = acc
        pkg = params[st["pkg_off"]:st["pkg_off"] + st["pkg_floats"]]
        for op in p.epis[st["epi_begin"]:st["epi_begin"] + st["n_epi"]]:
            k, dp, dt, flags = op["kind"], op["np"] * 8, op["dt"], op["flags"]
            gb = lambda j: pkg[op["off1"] * 4 + j * dp: op["off1"] * 4 + (j + 1) * dp][None, :]
            if k == T.OP_RAW_IN:
                t = torch.zeros(B, dp)
                t[:, :dt] = x
                emit(t, dp)
                continue
            if k == T.OP_RAW_S:
                emit(skips[op["slot"]][:, :dp].clone(), dp)
                continue
            v = regions[op["region"]][:, :dp].clone()
            if (flags & T.F_TIME) and images is None:
                v = v + table[t_idx][:, op["tt_src4"] * 4: op["tt_src4"] * 4 + dp]
            if flags & T.F_PUSH:
                skips[op["slot"]] = v.clone()
                skip_stats[op["slot"]] = moments(v, dt)
            if k == T.OP_OUT:
                out = v[:, :dt].clone()
            elif k == T.OP_RAW_T:
                emit(v, dp)                # pad columns are exact zeros
            elif k == T.OP_LN:
                s1, s2, shift = moments(v, dt)
                md = s1 / dt
                rstd = 1.0 / torch.sqrt((s2 / dt - md * md).clamp_min(0) + 1e-5)
                ln_emit(v, dp, shift + md, rstd, gb(0), gb(1))
                if (flags & T.F_COND):
                    emit(cpad, cpad.shape[1])
            elif k == T.OP_CATLN:   # LayerNorm over cat(x, skip); operands: skip part first, x part second
                s1, s2, shift = moments(v, dt)
                a1, a2, ashift = skip_stats[op["slot"]]
                mx, ms = shift + s1 / dt, ashift + a1 / dt
                m2 = (s2 - s1 * s1 / dt) + (a2 - a1 * a1 / dt) + (mx - ms) ** 2 * (0.5 * dt)
                mean = 0.5 * (mx + ms)
                rstd = 1.0 / torch.sqrt((m2 * (0.5 / dt)).clamp_min(0) + 1e-5)
                ln_emit(skips[op["slot"]], dp, mean, rstd, gb(2), gb(3))
                ln_emit(v, dp, mean, rstd, gb(0), gb(1))
            else:
                raise ValueError(k)
    assert not queue, f"{len(queue)} operand chunks were emitted but never consumed"
    return out
