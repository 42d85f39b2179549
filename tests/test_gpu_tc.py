"""GPU tests of the tcgen05 building blocks (one 128-row GEMM tile through the C-ABI hook)."""
import ctypes as C

import pytest
import torch

from diffsg_b200 import _lib

from conftest import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def pack_w_image(W: torch.Tensor, chunk_k: int = 64) -> torch.Tensor:
    """[N, K] fp16 -> concatenated per-K-chunk core-matrix images [n/8][k/8][n%8][8]."""
    N, K = W.shape
    parts = []
    for k0 in range(0, K, chunk_k):
        kw = min(chunk_k, K - k0)
        parts.append(W[:, k0:k0 + kw].reshape(N // 8, 8, kw // 8, 8).permute(0, 2, 1, 3).contiguous().reshape(-1))
    return torch.cat(parts).contiguous()


def run_gemm(A, W, nterms, layout=0, lbo=128, swap=0):
    lib = _lib.load()
    K, N = A.shape[1], W.shape[0]
    Wh = W.half()
    Wl = (W - Wh.float()).half()
    ih, il = pack_w_image(Wh).to(DEV), pack_w_image(Wl).to(DEV)
    Cd = torch.full((128, N), float("nan"), device=DEV)
    Ad = A.to(DEV).contiguous()
    _lib.check(lib.diffsg_debug_tc_gemm(Ad.data_ptr(), ih.data_ptr(), il.data_ptr(), Cd.data_ptr(), K, N, nterms,
                                        layout, lbo, swap, _lib.stream_ptr()), "diffsg_debug_tc_gemm")
    torch.cuda.synchronize()
    return Cd.cpu()


@pytest.mark.parametrize("K,N", [(16, 16), (64, 128), (128, 128), (256, 128), (80, 128), (128, 80), (32, 16), (192, 64)])
def test_tc_gemm_tile(K, N):
    g = torch.Generator().manual_seed(K * 1000 + N)
    A = torch.randn(128, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    want = A.double() @ W.double().t()
    e1 = rel_l2(run_gemm(A, W, 1), want)
    e2 = rel_l2(run_gemm(A, W, 2), want)
    e3 = rel_l2(run_gemm(A, W, 3), want)
    print(f"K={K} N={N}: rel-L2 fp16x1 {e1:.2e}  x2 {e2:.2e}  x3 {e3:.2e}")
    assert e1 < 2e-3 and e2 < 5e-4 and e3 < 2e-6


def test_two_programs_on_two_streams():
    """The stage program lives in __constant__ memory, one program resident per device: two plans used
    concurrently from two streams must still see their own program (the upload waits for the last kernel of the
    program it replaces)."""
    import diffsg_b200 as D
    from conftest import standin_model
    models = []
    for name in ("nu_like", "co"):
        ddpm, cfg = standin_model(name)
        ddpm = ddpm.to(DEV)
        ddpm.model.precision = "fp16x2"
        n = 40000                                    # 313 tiles: more than one wave, so launches overlap in time
        g = torch.Generator().manual_seed(7)
        x = torch.randn(n, cfg["input_dim"], generator=g).to(DEV)
        c = torch.rand(n, cfg["cond_dim"], generator=g).to(DEV)
        t = (torch.randint(0, 20, (1, n), generator=g) / 20).to(DEV)
        m = torch.ones(n, 1, device=DEV)
        with torch.no_grad():
            want = ddpm.model(x, t, c, m).clone()
        models.append((ddpm, x, t, c, m, want))
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(device=DEV) for _ in models]
    outs = [[] for _ in models]
    with torch.no_grad():
        for rep in range(6):
            for i, (ddpm, x, t, c, m, _) in enumerate(models):
                with torch.cuda.stream(streams[i]):
                    outs[i].append(ddpm.model(x, t, c, m))
    torch.cuda.synchronize()
    for i, (_, _, _, _, _, want) in enumerate(models):
        for o in outs[i]:
            assert torch.equal(o, want)
