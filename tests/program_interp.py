"""Test-only CPU interpreter of the op program emitted by diffsg_b200.packer (the same
semantics the CUDA kernels implement), used to check the lowering + packing host logic
against the oracle without a GPU."""
import torch

from diffsg_b200 import _lib


def run_program(prog, blob, table, x, t_idx, cond, mask):
    B = x.shape[0]
    W = 256
    bufs = [torch.zeros(B, W) for _ in range(4)]
    c = cond * mask
    bufs.append(c * torch.sigmoid(c))
    skips = {}
    bufs[prog.in_buf][:, :x.shape[1]] = x
    for o in prog.ops:
        k = o["kind"]
        if k == _lib.OP_GEMM:
            K, N, ldw = o["K"], o["N"], o["ldw"]
            Wm = blob[o["w_off"]:o["w_off"] + K * ldw].reshape(K, ldw)[:, :N]
            out = bufs[o["src"]][:, :K] @ Wm
            if not (o["flags"] & _lib.F_NOBIAS):
                out = out + blob[o["b_off"]:o["b_off"] + N]
            if o["flags"] & _lib.F_TIME:
                out = out + table[t_idx][:, o["t_off"]:o["t_off"] + N]
            dst = bufs[o["dst"]]
            if o["flags"] & _lib.F_ACC:
                out = out + dst[:, o["dcol"]:o["dcol"] + N]
            dst[:, o["dcol"]:o["dcol"] + N] = out
        elif k == _lib.OP_LNSW:
            D = o["N"]
            v = torch.nn.functional.layer_norm(bufs[o["src"]][:, :D], (D,), blob[o["w_off"]:o["w_off"] + D],
                                               blob[o["b_off"]:o["b_off"] + D], 1e-5)
            bufs[o["dst"]][:, :D] = v * torch.sigmoid(v)
        elif k == _lib.OP_PUSH:
            skips[o["dcol"]] = bufs[o["src"]][:, :o["N"]].clone()
        elif k == _lib.OP_POP:
            bufs[o["dst"]][:, o["dcol"]:o["dcol"] + o["N"]] = skips[o["K"]]
        else:
            raise ValueError(k)
    return bufs[prog.out_buf][:, :x.shape[1]].clone()
