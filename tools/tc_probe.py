import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from test_gpu_tc import run_gemm
from conftest import rel_l2
g = torch.Generator().manual_seed(0)
for (K, N) in [(16, 16), (64, 64), (128, 128)]:
    A = torch.randn(128, K, generator=g); W = torch.randn(N, K, generator=g) / K ** 0.5
    want = A.double() @ W.double().t()
    for swap in (0, 1):
        for nt in (1, 2, 3):
            try:
                got = run_gemm(A, W, nt, 0, 128, swap)
                print(f"K={K} N={N} swap={swap} nterms={nt}: rel-L2 {rel_l2(got, want):.3e} nan={int(torch.isnan(got).sum())}", flush=True)
            except Exception as e:
                print("ERR", K, N, swap, nt, e, flush=True); raise
