"""The tcgen05 training kernels (csrc/train_tc.cu, C-ABI diffsg_tlin_*) against a float64 torch restatement of
the same node: y = swish(LN(cat(x0, x1))) . W^T + b + cat(z0, z1) . W2^T + b2 + add + gadd[gidx], and every
gradient of it (reference semantics: ddpm_opt/UNetCF.py:83-95 under autograd).  bf16 hi+lo operands carry 16
significant bits: the gate is 5e-5 rel-L2 per tensor (measured ~5e-6)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
device = torch.device("cuda:0")


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _ref_node(x0, x1, gamma, beta, w, b, z0, z1, w2, b2, add, gadd, gidx):
    d = lambda t: None if t is None else t.detach().double().requires_grad_(t.requires_grad)     # noqa: E731
    T = dict(x0=d(x0), x1=d(x1), gamma=d(gamma), beta=d(beta), w=d(w), b=d(b), z0=d(z0), z1=d(z1), w2=d(w2), b2=d(b2),
             add=d(add), gadd=d(gadd))
    a = T["x0"] if x1 is None else torch.cat((T["x0"], T["x1"]), 1)
    if gamma is not None:
        a = F.silu(F.layer_norm(a, (a.shape[1],), T["gamma"], T["beta"], 1e-5))
    y = F.linear(a, T["w"], T["b"])
    if w2 is not None:
        z = T["z0"] if z1 is None else torch.cat((T["z0"], T["z1"]), 1)
        y = y + F.linear(z, T["w2"], T["b2"])
    if add is not None:
        y = y + T["add"]
    if gadd is not None:
        y = y + T["gadd"][gidx]
    return y, T


CASES = [
    # B,   k0,  k1, N,   ln,    z0k, z1k, add,   T
    (300, 128, 0, 128, True, 0, 0, True, 0),
    (128, 128, 128, 128, True, 128, 128, False, 0),       # UpBlock lin1 / lin3 + shortcut on the cat
    (257, 64, 0, 64, True, 80, 0, False, 0),              # lin2 + cond_emb
    (129, 32, 32, 16, True, 0, 0, False, 20),             # gathered time term
    (77, 8, 0, 8, True, 0, 0, True, 0),
    (40, 16, 0, 8, True, 16, 0, False, 20),
    (20, 512, 0, 512, False, 0, 0, False, 0),             # time MLP (rows = T)
    (20, 128, 0, 512, False, 0, 0, False, 0),
    (500, 80, 0, 128, False, 0, 0, False, 0),             # feature_proj
    (131, 3, 0, 24, False, 0, 0, False, 0),               # odd widths (MSR 3c / NU input)
    (66, 24, 0, 3, True, 5, 0, True, 7),
    (1, 128, 0, 80, True, 0, 0, False, 1),
    (1000, 256, 0, 96, False, 0, 0, True, 0),
    (640, 48, 0, 200, False, 0, 0, False, 0),             # N tiles 128 + 72
]


@pytest.mark.parametrize("case", CASES, ids=[str(c) for c in CASES])
def test_fused_linear_forward_and_backward(case):
    from diffsg_b200.train import _FusedLinear
    B, k0, k1, N, ln, z0k, z1k, has_add, T = case
    g = torch.Generator(device="cpu").manual_seed(hash(case) % (2 ** 31))
    r = lambda *s, scale=1.0, grad=True: (torch.randn(*s, generator=g) * scale).to(device).requires_grad_(grad)      # noqa: E731
    K = k0 + k1
    x0 = r(B, k0, scale=2.0)
    x1 = r(B, k1) if k1 else None
    gamma = r(K) if ln else None
    beta = r(K, scale=0.3) if ln else None
    w, b = r(N, K, scale=K ** -0.5), r(N, scale=0.1)
    z0 = r(B, z0k) if z0k else None
    z1 = r(B, z1k) if z1k else None
    w2 = r(N, z0k + z1k, scale=(z0k + z1k) ** -0.5) if z0k else None
    b2 = r(N, scale=0.1) if z0k else None
    add = r(B, N) if has_add else None
    gadd = r(T, N) if T else None
    gidx = torch.randint(0, T, (B,), generator=g).to(device) if T else None
    args = (x0, x1, gamma, beta, w, b, z0, z1, w2, b2, add, gadd, gidx)
    y = _FusedLinear.apply(*args)
    want, R = _ref_node(*args)
    assert y.shape == want.shape
    assert rel_l2(y, want) < 5e-5, rel_l2(y, want)
    dy = (torch.randn(B, N, generator=g) * 1e-3).to(device)          # small gradients: fp16 would lose them
    y.backward(dy)
    want.backward(dy.double())
    names = ("x0", "x1", "gamma", "beta", "w", "b", "z0", "z1", "w2", "b2", "add", "gadd")
    for name, t in zip(names, args[:12]):
        if t is None:
            continue
        assert t.grad is not None, name
        err = rel_l2(t.grad, R[name].grad)
        assert err < 5e-5, (name, err)


def test_parameter_gradients_accumulate_in_place():
    """wgrad adds into an existing .grad (autograd's += semantics) without reallocating it."""
    from diffsg_b200.train import fused_linear
    lin = torch.nn.Linear(32, 16).to(device)
    norm = torch.nn.LayerNorm(32).to(device)
    x = torch.randn(200, 32, device=device)
    for p in list(lin.parameters()) + list(norm.parameters()):
        p.grad = torch.ones_like(p)
    ptrs = [p.grad.data_ptr() for p in lin.parameters()]
    fused_linear(x, lin, norm=norm).sum().backward()
    ref_lin, ref_norm = torch.nn.Linear(32, 16).to(device), torch.nn.LayerNorm(32).to(device)
    ref_lin.load_state_dict(lin.state_dict())
    ref_norm.load_state_dict(norm.state_dict())
    ref_lin(F.silu(ref_norm(x))).sum().backward()
    for p, q in zip(list(lin.parameters()) + list(norm.parameters()), list(ref_lin.parameters()) + list(ref_norm.parameters())):
        assert torch.allclose(p.grad, q.grad + 1.0, rtol=1e-4, atol=1e-4)
    assert ptrs == [p.grad.data_ptr() for p in lin.parameters()]


def test_empty_batch_launches_nothing():
    from diffsg_b200 import _lib
    from diffsg_b200.train import fused_linear
    lin = torch.nn.Linear(8, 8).to(device)
    before = _lib.launch_count()
    y = fused_linear(torch.zeros(0, 8, device=device), lin)
    assert y.shape == (0, 8) and _lib.launch_count() == before


@pytest.mark.parametrize("ln", [False, True])
def test_dgrad_addend_through_the_c_abi(ln):
    """diffsg_tlin_dgrad with `dres` (an addend on dx, e.g. the gradient arriving over a residual connection) and a cat
    output split, called through the C-ABI directly."""
    import ctypes as C
    from diffsg_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(5)
    B, N, k0, k1 = 200, 48, 64, 32
    K = k0 + k1
    r = lambda *s_: torch.randn(*s_, generator=g).to(device)      # noqa: E731
    dy, w, x, res = r(B, N), r(N, K) * K ** -0.5, r(B, K), r(B, K)
    gamma, beta = r(K), r(K) * 0.3
    xd = x.double().requires_grad_(True)
    a = F.silu(F.layer_norm(xd, (K,), gamma.double(), beta.double(), 1e-5)) if ln else xd
    (F.linear(a, w.double()) * dy.double()).sum().backward()
    want = xd.grad + res.double()
    mean = x.mean(1).contiguous()
    rstd = (x.var(1, unbiased=False) + 1e-5).rsqrt().contiguous()
    dx0, dx1 = torch.empty(B, k0, device=device), torch.empty(B, k1, device=device)
    r0, r1 = res[:, :k0].contiguous(), res[:, k0:].contiguous()
    x0, x1 = x[:, :k0].contiguous(), x[:, k0:].contiguous()
    dg, db = torch.zeros(K, device=device), torch.zeros(K, device=device)
    args = _lib.TlinDgradArgs(dy=dy.data_ptr(), w=w.data_ptr(), x=_lib.Mat(x0.data_ptr(), x1.data_ptr(), k0, k1) if ln else _lib.Mat(None, None, 0, 0),
                              gamma=gamma.data_ptr() if ln else None, beta=beta.data_ptr() if ln else None,
                              mean=mean.data_ptr() if ln else None, rstd=rstd.data_ptr() if ln else None,
                              dres=_lib.Mat(r0.data_ptr(), r1.data_ptr(), k0, k1), dx=_lib.Mat(dx0.data_ptr(), dx1.data_ptr(), k0, k1),
                              dgamma=dg.data_ptr() if ln else None, dbeta=db.data_ptr() if ln else None, B=B, N=N, K=K)
    _lib.check(lib.diffsg_tlin_dgrad(C.byref(args), _lib.stream_ptr()), "diffsg_tlin_dgrad")
    torch.cuda.synchronize()
    assert rel_l2(torch.cat((dx0, dx1), 1), want) < 5e-5
