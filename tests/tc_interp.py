"""Test-only CPU interpreter of the tensor-core stage program (diffsg_b200.tc_packer): same
dataflow as diffsg_b200/csrc/unet_tc.cuh (TMEM regions, per-row vector, operand chunk queue,
per-stage parameter packages), evaluated in fp32 (optionally with the fp16 operand rounding of
the real engine)."""
import torch

from diffsg_b200 import tc_packer as T


def run_tc_program(p, w_hi, w_lo, params, table, x, t_idx, cond, mask, emulate_fp16=False):
    B = x.shape[0]
    regions = [torch.zeros(B, 128), torch.zeros(B, 128)]
    skips = {}
    c = cond * mask
    cpad = torch.zeros(B, T.pad16(p.cond_dim))
    cpad[:, :p.cond_dim] = c * torch.sigmoid(c)
    queue = []                       # emitted A chunks, FIFO
    v = torch.zeros(B, 128)
    stats = dict(cnt=0.0, mean=torch.zeros(B), m2=torch.zeros(B), rstd=torch.ones(B))
    out = None

    def split(a):
        if not emulate_fp16:
            return a
        hi = a.half().float()
        return hi + (a - hi).half().float()

    def emit(vec, dp):
        for k0 in range(0, dp, T.CHUNK_K):
            queue.append(split(vec[:, k0:min(dp, k0 + T.CHUNK_K)].clone()))

    def do_stats(dt, flags):
        if flags & T.STATS_RESET:
            stats.update(cnt=0.0, mean=torch.zeros(B), m2=torch.zeros(B))
        m = v[:, :dt].mean(dim=1)
        q = ((v[:, :dt] - m[:, None]) ** 2).sum(dim=1)
        tot = stats["cnt"] + dt
        delta = m - stats["mean"]
        stats["mean"] = stats["mean"] + delta * (dt / tot)
        stats["m2"] = stats["m2"] + q + delta * delta * (stats["cnt"] * dt / tot)
        stats["cnt"] = tot
        if flags & T.STATS_FINISH:
            stats["rstd"] = 1.0 / torch.sqrt(stats["m2"] / tot + 1e-5)

    def do_emit_ln(pkg, dp, dt, og, ob):
        g, b = pkg[:, og * 4:og * 4 + dp], pkg[:, ob * 4:ob * 4 + dp]
        t = (v[:, :dp] - stats["mean"][:, None]) * stats["rstd"][:, None] * g + b
        t = t * torch.sigmoid(t)
        t[:, dt:] = 0
        emit(t, dp)

    for st in p.stages:
        if st["has_gemm"]:
            N = st["n16"] * 16
            acc = regions[st["region"]][:, :N].clone() if st["accumulate"] else torch.zeros(B, N)
            for ch in p.chunks[st["chunk_begin"]:st["chunk_begin"] + st["n_chunks"]]:
                kw = ch["kw"]
                a = queue.pop(0)
                assert a.shape[1] == kw, (a.shape, kw)
                off = ch["w_off16"] * 8
                img = w_hi[off:off + N * kw].float()
                if p.nterms >= 3:
                    img = img + w_lo[off:off + N * kw].float()
                W = img.reshape(N // 8, kw // 8, 8, 8).permute(0, 2, 1, 3).reshape(N, kw)
                acc = acc + a @ W.t()
            regions[st["region"]][:, :N] = acc
        # per-row package = [time slice of the row's table entry | static part]
        static = params[st["pkg_off"]:st["pkg_off"] + st["pkg_floats"]][None, :].expand(B, -1)
        if st["tt_src"] is not None:
            tt = table[t_idx][:, st["tt_src"]:st["tt_src"] + st["tt_floats"]]
            pkg = torch.cat((tt, static), dim=1)
        else:
            pkg = static
        for op in p.epis[st["epi_begin"]:st["epi_begin"] + st["n_epi"]]:
            k, dp, dt = op["kind"], op["np"] * 8, op["dt"]
            bias = lambda o4: pkg[:, o4 * 4:o4 * 4 + dp]
            if k == T.OP_RAW_IN:
                t = torch.zeros(B, dp)
                t[:, :dt] = x
                emit(t, dp)
            elif k == T.OP_RAW_S:
                t = skips[op["slot"]][:, :dp].clone()
                t[:, dt:] = 0
                emit(t, dp)
            elif k in (T.OP_RAW_T, T.OP_LN, T.OP_OUT, T.OP_CATLN):
                v[:, :dp] = regions[op["region"]][:, :dp] + bias(op["off0"])
                if op["flags"] & T.F_PUSH:
                    skips[op["slot"]] = v[:, :dp].clone()
                if k == T.OP_OUT:
                    out = v[:, :dt].clone()
                elif k == T.OP_RAW_T:
                    t = v[:, :dp].clone()
                    t[:, dt:] = 0
                    emit(t, dp)
                elif k == T.OP_LN:
                    do_stats(dt, T.STATS_RESET | T.STATS_FINISH)
                    do_emit_ln(pkg, dp, dt, op["off1"], op["off2"])
                    if op["flags"] & T.F_COND:
                        emit(cpad, cpad.shape[1])
                else:   # CATLN: LayerNorm over cat(x, skip); operands: skip part first, x part second
                    dp4 = dp // 4
                    xv = v[:, :dp].clone()
                    do_stats(dt, T.STATS_RESET)
                    v[:, :dp] = skips[op["slot"]][:, :dp]
                    do_stats(dt, T.STATS_FINISH)
                    do_emit_ln(pkg, dp, dt, op["off1"] + 2 * dp4, op["off1"] + 3 * dp4)
                    v[:, :dp] = xv
                    do_emit_ln(pkg, dp, dt, op["off1"], op["off1"] + dp4)
            else:
                raise ValueError(k)
    assert not queue, f"{len(queue)} operand chunks were emitted but never consumed"
    return out
