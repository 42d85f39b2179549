"""CPU-side tests: oracle vs golden vectors, host logic (schedule, state_dict layout, program
lowering + packing + hoisting), C-ABI library surface.  No GPU needed."""
import ctypes
import re
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

import diffsg_b200 as D
from diffsg_b200 import _lib
from diffsg_b200.packer import lower, pack_params, time_table
from oracle import ddpm_oracle as O
from oracle.standin import CONFIGS, make_state_dict

from conftest import ROOT, load_golden, nu_checkpoint_model, rel_l2, standin_model
from program_interp import run_program

T = 20


def test_schedule_matches_golden():
    g = load_golden("schedule_T20.npz")
    betas = D.generate_cosine_schedule(T)
    assert np.array_equal(betas, g["betas64"])
    assert np.array_equal(O.cosine_betas(T), g["betas64"])
    assert abs(betas[0] - 0.00799272) < 1e-8 and betas[19] == 0.84
    ddpm, _ = standin_model("attn")
    for k in ("betas", "alphas", "alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
              "reciprocal_sqrt_alphas", "remove_noise_coeff", "sqrt_betas"):
        assert np.array_equal(getattr(ddpm, k).numpy(), g[k]), k


@pytest.mark.parametrize("name", list(CONFIGS))
def test_state_dict_layout_matches_reference(name, manifest):
    ddpm, _ = standin_model(name)
    sd = ddpm.state_dict()
    ref = manifest[name]
    assert list(sd.keys()) == sorted(ref, key=list(ref).index) or set(sd.keys()) == set(ref)
    assert set(sd.keys()) == set(ref)
    for k, v in sd.items():
        assert list(v.shape) == ref[k], k
    assert sd["ema.n_averaged"].dtype == torch.int64


def test_nu_checkpoint_keys(manifest):
    ddpm = nu_checkpoint_model()
    ref = manifest["nu_ckpt"]
    sd = ddpm.state_dict()
    assert len(ref) == 805 and set(sd.keys()) == set(ref)
    assert all(list(sd[k].shape) == ref[k] for k in ref)
    cfg = D.infer_config_from_state_dict(sd)
    assert cfg == dict(input_dim=5, proj_dim=32, cond_dim=6, dims=(32, 16, 8), is_attn=(False,) * 3,
                       middle_attn=False, n_blocks=2)


def test_infer_config_roundtrip():
    for name, (_, cfg) in CONFIGS.items():
        m = D.UNet1D(**cfg)
        got = D.infer_config_from_state_dict(m.state_dict(), prefix="")
        assert got == {**cfg, "dims": tuple(cfg["dims"]), "is_attn": tuple(cfg["is_attn"])}, name


@pytest.mark.parametrize("name", list(CONFIGS))
def test_oracle_reproduces_reference_goldens(name):
    """The committed eps / sampler outputs came from the unmodified reference; the oracle must
    reproduce them bit-for-bit on this machine too."""
    g = load_golden(f"standin_{name}.npz")
    _, cfg = CONFIGS[name]
    shapes = {k: v.shape for k, v in D.UNet1D(**cfg).state_dict().items()}
    sd = {"model." + k: v for k, v in make_state_dict(shapes, 1234).items()}
    sd.update(O.ddpm_buffers(1.0 - O.cosine_betas(T)))
    x, cond, ts, mask = (torch.tensor(g[k]) for k in ("x", "cond", "ts", "mask"))
    eps = O.unet_forward(sd, x, ts / T, cond, mask)
    assert rel_l2(eps, g["eps"]) < 1e-6
    if name in ("attn", "nu_like"):
        M = cfg["input_dim"]
        B = x.shape[0]
        steps = [torch.tensor(n).reshape(B, 1, M) for n in g["noise"]]
        y0 = O.sample(sd, T, cond, 3.0, torch.tensor(g["y_T"]).reshape(B, 1, M), steps)
        assert rel_l2(y0, g["y0_omega3"]) < 1e-5


def test_oracle_nu_checkpoint_trace():
    g = load_golden("nu_trace.npz")
    ck = {k: torch.tensor(v) for k, v in load_golden("nu_ckpt.npz").items()}
    B = 64
    cond = torch.tensor(g["cond"][:B])
    for step in (0, 7, 19):
        i = T - 1 - step
        t = torch.full((1, B), i) / T
        y = torch.tensor(g["y_in"][step][:B])
        e1 = O.unet_forward(ck, y, t, cond, torch.ones(B, 1))
        e0 = O.unet_forward(ck, y, t, cond, torch.zeros(B, 1))
        assert rel_l2(e1, g["eps_1"][step][:B]) < 1e-6 and rel_l2(e0, g["eps_0"][step][:B]) < 1e-6


@pytest.mark.parametrize("name", list(CONFIGS))
def test_lowered_program_matches_oracle(name):
    """packer: topology -> ops, weight transposition/padding, cond-bias folding, time hoisting."""
    g = load_golden(f"standin_{name}.npz")
    ddpm, cfg = standin_model(name)
    prog = lower(ddpm.model)
    blob = pack_params(prog, "cpu")
    ts = torch.tensor(g["ts"]).reshape(-1)
    table = time_table(ddpm.model, prog, torch.arange(T) / T)
    x, cond, mask = (torch.tensor(g[k]) for k in ("x", "cond", "mask"))
    eps = run_program(prog, blob, table, x, ts, cond, mask)
    assert rel_l2(eps, g["eps"]) < 2e-6
    x_macs, c_macs = prog.gemm_macs()
    expect = {"msr3c": (546688, 3768), "msr80c": (566400, 100480), "co": (329024, 11736), "nu_like": (60864, 2736)}
    if name in expect:
        assert (x_macs, c_macs) == expect[name]


def test_program_buffers_never_alias():
    for name, (_, cfg) in CONFIGS.items():
        prog = lower(D.UNet1D(**cfg))
        for o in prog.ops:
            if o["kind"] in (_lib.OP_GEMM, _lib.OP_LNSW):
                assert o["src"] != o["dst"], (name, o)
            if o["kind"] == _lib.OP_GEMM:
                assert o["w_off"] % 4 == 0 and o["ldw"] % 4 == 0 and o["ldw"] >= o["N"]
        assert prog.max_width <= 256


def test_philox_known_answers():
    """Random123 known-answer vectors for Philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = O.philox4x32_10(np.array([ctr], dtype=np.uint32), key)[0]
        assert tuple(int(v) for v in got) == want
    z = O.philox_normal(4096, 80, 7, seed=42)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01


def test_objective_oracles_match_goldens():
    g = load_golden("msr_data.npz")
    dec = O.msr_decode(torch.tensor(g["y_rand"]))
    assert torch.equal(dec, torch.tensor(g["dec_rand"]))
    lo, hi, W = g["scaler"]
    gains = torch.tensor(g["X_test"]) * (hi - lo) + lo
    assert abs(float(O.msr_rate(torch.tensor(g["Y_test"]), gains.float()).mean()) - 7.533) < 5e-3
    c = load_golden("co_data.npz")
    lo, hi = c["scaler"]
    Xs = (torch.tensor(c["X_test"]) * (hi - lo) + lo).float()
    cost = O.co_cost(Xs, torch.tensor(c["Y_test"]))
    assert rel_l2(cost, c["true_cost"]) < 1e-6 and abs(float(cost.mean()) - 2.0587) < 5e-3
    assert rel_l2(O.co_cost(Xs, O.co_decode(torch.tensor(c["y_rand"]))), c["pred_cost_rand"]) < 1e-6
    n = load_golden("nu_objective.npz")
    d = load_golden("nu_data.npz")
    dec = O.nu_decode(torch.tensor(n["y0_test"]), 400, 400, 18.0)
    rate = O.nu_rate(dec, torch.tensor(d["X_test"]) * 400.0)
    assert rel_l2(rate, n["pred_rate_test"]) < 1e-6
    assert rel_l2(rate[:200], n["rate_calc_ref_first200"]) < 1e-6  # reference rate_calc itself
    assert 0.88 < float(n["less_ratio_test"]) < 0.93 and 0.89 < float(n["less_ratio_ood"]) < 0.94


def test_loaders_handle_reference_names(tmp_path):
    assert D.msr.parse_scalar_from_name("../datasets/3c_10w_10000samples.csv", "w") == 10.0
    assert D.msr.parse_scalar_from_name("3c_20w_2000samples_ood.csv", "w") == 20.0
    assert D.msr.parse_scalar_from_name("x/3u_18mW_10000samples.csv", "mw") == 18.0
    assert D.msr.parse_scalar_from_name("3u_30mW_1000samples_ood.csv", "mw") == 30.0
    rng = np.random.default_rng(0)
    rows = np.concatenate([rng.uniform(0.5, 2.5, (50, 3)), rng.uniform(5, 8, (50, 1)), rng.uniform(0, 10, (50, 3))], 1)
    p = tmp_path / "3c_10w_50samples.csv"
    np.savetxt(p, rows, delimiter=",")
    Xtr, Ytr, Xte, Yte, cfg = D.msr.msr_data_load(str(p))
    assert Xtr.shape == (35, 3) and Xte.shape == (15, 3) and cfg["W"] == 10.0 and cfg["M"] == 3
    assert Xtr.min() >= 0 and Xtr.max() <= 1


def test_no_cpu_fallback():
    ddpm, cfg = standin_model("attn")
    with pytest.raises(_lib.DiffsgError):
        ddpm.sample(torch.rand(4, cfg["cond_dim"]), 1.0)
    with pytest.raises(_lib.DiffsgError):
        with torch.no_grad():
            ddpm.model(torch.rand(4, cfg["input_dim"]), torch.zeros(1, 4), torch.rand(4, cfg["cond_dim"]), torch.ones(4, 1))
    with pytest.raises(_lib.DiffsgError):
        D.objectives.co_cost(torch.rand(4, 9), torch.rand(4, 3))


def test_reference_aliases():
    import sys
    saved = {k: v for k, v in sys.modules.items() if k == "ddpm_opt" or k.startswith("ddpm_opt.")}
    for k in saved:
        del sys.modules[k]
    try:
        D.install_reference_aliases()
        from ddpm_opt.UNetCF import UNet1D
        from ddpm_opt.classifier_free_NU import DDPM, custom_decoder, nu_data_load, rate_calc  # noqa: F401
        from ddpm_opt.classifier_free_MSR import DDPM as M2, msr_data_load  # noqa: F401
        from ddpm_opt.classifier_free_CO import DDPM as C2, co_data_load, cost_calc  # noqa: F401
        from ddpm_opt.diffusion import generate_cosine_schedule, init_weights  # noqa: F401
        from ddpm_opt.ema import ExponentialMovingAverage  # noqa: F401
        assert UNet1D is D.UNet1D and DDPM is D.nu.DDPM
    finally:
        for k in [k for k in sys.modules if k == "ddpm_opt" or k.startswith("ddpm_opt.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_c_abi_exports_every_declared_symbol():
    """The shared library loads and exports exactly what include/diffsg_b200.h declares."""
    header = (ROOT / "include" / "diffsg_b200.h").read_text()
    declared = set(re.findall(r"\b(diffsg_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    _lib.build_library()
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().diffsg_abi_version() == _lib.ABI_VERSION
    assert ctypes.sizeof(_lib.Op) == 48 and ctypes.sizeof(_lib.Cfg) == 64 and ctypes.sizeof(_lib.SampleArgs) == 96
    assert (ctypes.sizeof(_lib.Mat), ctypes.sizeof(_lib.TlinFwdArgs), ctypes.sizeof(_lib.TlinDgradArgs),
            ctypes.sizeof(_lib.TlinWgradArgs)) == (24, 160, 152, 120)        # static_assert'ed in include/diffsg_b200.h


@pytest.mark.parametrize("name", ["nu_like", "msr3c", "msr80c", "co", "attn"])
def test_tc_program_matches_oracle(name):
    """tc_packer: stage/chunk/epilogue lowering, fp16 weight images (x3 -> ~fp32), biases as K = 16 chunks on a
    constant ones tile (static images + per-step time images), cat-free UpBlocks with the skip moments stored at
    push time, merged lin3+shortcut GEMM groups — interpreted on the CPU."""
    from diffsg_b200 import tc_packer
    from tc_interp import run_tc_program
    g = load_golden(f"standin_{name}.npz")
    ddpm, cfg = standin_model(name)
    prog = tc_packer.lower_tc(ddpm.model, nterms=3)
    hi, lo, params = tc_packer.pack_tc_weights(prog, "cpu")
    table = tc_packer.time_table_tc(ddpm.model, prog, torch.arange(T) / T)
    x, cond, mask = (torch.tensor(g[k]) for k in ("x", "cond", "mask"))
    ts = torch.tensor(g["ts"]).reshape(-1)
    eps = run_tc_program(prog, hi, lo, params, table, x, ts, cond, mask)          # forward mode: fp32 table rows
    assert rel_l2(eps, g["eps"]) < 5e-6
    images = tc_packer.time_images(prog, table)                                    # sampler mode: time-bias chunk images
    assert images.shape == (T, prog.img_stride // 2) and prog.img_stride % 16 == 0
    for step in sorted(set(ts.tolist()))[:3]:
        sel = ts == step
        eps_s = run_tc_program(prog, hi, lo, params, table, x[sel], ts[sel], cond[sel], mask[sel], images=images)
        assert rel_l2(eps_s, g["eps"][sel.numpy()]) < 5e-6
    st, ch, ep = prog.arrays()
    assert ch.dtype.itemsize == 8 and ep.dtype.itemsize == 8
    assert len(st) <= 256 and len(ch) <= 640 and len(ep) <= 512 and st.dtype.itemsize == 16      # kMax* of unet_tc.cuh
    # every GEMM group carries one bias image; the three fp16 terms reproduce the fp32 bias to 2^-30
    assert all(sdict["has_bias"] for sdict in prog.stages if sdict["has_gemm"])
    assert sum(sdict["time_bias"] for sdict in prog.stages) == len(prog.time_blocks) > 0
    b = torch.randn(3, 40) * torch.tensor([1e-6, 1.0, 300.0])[:, None]
    img = tc_packer.bias_image(b, 48).float().reshape(3, 6, 2, 8, 8)
    back = img[:, :, 0, :, :3].sum(-1).reshape(3, 48)
    assert torch.all((back[:, :40] - b).abs() <= b.abs() * 2.0 ** -30 + 2.0 ** -25) and torch.all(back[:, 40:] == 0)
    assert torch.all(img[:, :, 1] == 0) and torch.all(img[:, :, 0, :, 3:] == 0)
    expect = {"msr3c": (546688, 3768), "msr80c": (566400, 100480), "co": (329024, 11736), "nu_like": (60864, 2736)}
    if name in expect:
        assert prog.gemm_macs() == expect[name]
    else:   # attention (UNetCF.py:123-157) = one accumulate stage per block whose raw operand is published deferred
        raw_t = [e for e in prog.epis if e["kind"] == tc_packer.OP_RAW_T and e["flags"] & tc_packer.F_DEFER]
        assert len(raw_t) == sum(isinstance(m, D.unet.AttentionBlock) for m in ddpm.model.modules()) > 0
    # fp16x2 mode: same program, weights rounded to fp16 once
    prog2 = tc_packer.lower_tc(ddpm.model, nterms=2)
    hi2, lo2, params2 = tc_packer.pack_tc_weights(prog2, "cpu")
    assert lo2 is None
    eps2 = run_tc_program(prog2, hi2, None, params2, table, x, ts, cond, mask, emulate_fp16=True)
    assert rel_l2(eps2, g["eps"]) < 1e-3


def test_tc_engine_rejects_unsupported_topologies():
    from diffsg_b200 import tc_packer
    wide = D.UNet1D(input_dim=4, proj_dim=256, cond_dim=4, dims=(64, 32), is_attn=(False, False), n_blocks=1)
    assert tc_packer.supported(wide) is not None
    odd = D.UNet1D(input_dim=4, proj_dim=24, cond_dim=4, dims=(12, 6), is_attn=(False, False), n_blocks=1)
    assert tc_packer.supported(odd) is None          # any width <= 128 (padded to multiples of 16)
    assert tc_packer.supported(D.UNet1D(input_dim=200, proj_dim=32, cond_dim=4, dims=(16, 8), is_attn=(False, False), n_blocks=1))


ODD_NETS = [dict(input_dim=4, proj_dim=24, cond_dim=5, dims=(12, 6), is_attn=(False, True), middle_attn=True, n_blocks=1),
            dict(input_dim=7, proj_dim=40, cond_dim=3, dims=(40, 20, 10), is_attn=(False,) * 3, middle_attn=False, n_blocks=2),
            dict(input_dim=3, proj_dim=100, cond_dim=9, dims=(72, 36), is_attn=(False,) * 2, middle_attn=False, n_blocks=1)]


@pytest.mark.parametrize("cfg", ODD_NETS, ids=lambda c: "x".join(map(str, (c["proj_dim"],) + tuple(c["dims"]))))
def test_tc_program_odd_widths_match_oracle(cfg):
    """Widths that are not multiples of 16 (nor powers of two): padded columns are exact zeros end to end."""
    from diffsg_b200 import tc_packer
    from oracle.standin import make_state_dict
    from tc_interp import run_tc_program
    model = D.UNet1D(**cfg)
    sd = make_state_dict({k: v.shape for k, v in model.state_dict().items()}, seed=7)
    model.load_state_dict(sd)
    B = 33
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, cfg["input_dim"], generator=g)
    c = torch.rand(B, cfg["cond_dim"], generator=g)
    ts = torch.randint(0, T, (1, B), generator=g)
    m = (torch.rand(B, 1, generator=g) > 0.3).float()
    want = O.unet_forward({"model." + k: v for k, v in sd.items()}, x, ts / T, c, m)
    for nterms, emulate, tol in ((3, False, 5e-6), (2, True, 1e-3)):
        prog = tc_packer.lower_tc(model, nterms=nterms)
        hi, lo, params = tc_packer.pack_tc_weights(prog, "cpu")
        table = tc_packer.time_table_tc(model, prog, torch.arange(T) / T)
        eps = run_tc_program(prog, hi, lo, params, table, x, ts.reshape(-1), c, m, emulate_fp16=emulate)
        assert rel_l2(eps, want) < tol


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver times beside ours) prints ONE JSON line with the
    contract's keys; it must run without a GPU."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-rows", "32"], capture_output=True, text=True, timeout=300, cwd=str(ROOT))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "solutions/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1 and d["gpu_launches"] == 0
    # the unmodified reference (baseline/_ref, copied by __graft_entry__.build()) when it is there, else the oracle port
    assert d["cpu_baseline"]["kind"] == ("reference" if (ROOT / "baseline" / "_ref" / "ddpm_opt").exists() else "port")
    assert d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    # both arms describe the workload with the same `config` (the GPU arm reports its engine outside it)
    import bench
    assert d["config"] == bench.sample_config(1 << 20) and d["metric"] == bench.METRIC


REF = Path("/root/reference")


@pytest.mark.skipif(not (REF / "datasets").exists(), reason="reference datasets not mounted here")
def test_loaders_match_reference_loader_goldens(tmp_path):
    """The three loaders on the bundled CSVs against tests/golden/{msr,nu,co}_data.npz, which oracle/make_golden.py
    produced with the reference's own loaders (MSR.py:159-184, NU.py:184-210, CO.py:158-200)."""
    f32 = lambda a: np.asarray(a).astype(np.float32)        # the fixtures were stored as float32
    g = load_golden("msr_data.npz")
    Xtr, Ytr, Xte, Yte, cfg = D.msr.msr_data_load(str(REF / "datasets" / "3c_10w_10000samples.csv"))
    assert np.array_equal(f32(Xte), g["X_test"]) and np.array_equal(f32(Yte), g["Y_test"]) and np.array_equal(f32(Xtr), g["X_train"])
    assert cfg["W"] == 10.0 and cfg["M"] == 3 and np.allclose([cfg["scaler_min"], cfg["scaler_max"]], g["scaler"][:2], rtol=0, atol=0)
    # the OOD file name defeats the reference's parser; ours reads it
    assert D.msr.msr_data_load(str(REF / "datasets" / "3c_20w_2000samples_ood.csv"))[4]["W"] == 20.0
    g = load_golden("nu_data.npz")
    Xtr, Ytr, Xte, Yte, Rte, cfg = D.nu.nu_data_load(str(REF / "datasets" / "3u_18mW_10000samples.csv"), 400, 400)
    assert np.array_equal(f32(Xte), g["X_test"]) and np.array_equal(f32(Yte), g["Y_test"]) and np.array_equal(f32(Rte), g["R_test"])
    assert np.array_equal(f32(Xtr[:1024]), g["X_train_head"]) and cfg["P_sum"] == 18.0 and cfg["K"] == 3
    _, _, Xo, Yo, _, ocfg = D.nu.nu_data_load(str(REF / "datasets" / "3u_30mW_1000samples_ood.csv"), 400, 400)
    assert ocfg["P_sum"] == 30.0 and np.array_equal(f32(Xo), g["X_ood"][-Xo.shape[0]:]) and np.array_equal(f32(Yo), g["Y_ood"][-Yo.shape[0]:])
    g = load_golden("co_data.npz")
    Xtr, Ytr, Xte, Yte, cfg = D.co.co_data_load(str(REF / "datasets" / "3nodes_2000samples_ood.csv"))
    assert np.allclose(f32(Xte), g["X_test"], rtol=0, atol=1e-7) and np.array_equal(f32(Yte), g["Y_test"])
    assert np.allclose([cfg["scaler_min"], cfg["scaler_max"]], g["scaler"])


def test_multistep_lr_matches_torch():
    from diffsg_b200.parallel import MultiStepLR

    class Opt:                                   # stand-in for FusedAdam's host side (set_lr only)
        lr = 0.005

        def set_lr(self, lr):
            self.lr = lr

    o = Opt()
    s = MultiStepLR(o, [15, 80, 150])
    p = torch.nn.Parameter(torch.zeros(1))
    topt = torch.optim.Adam([p], lr=0.005)
    ts = torch.optim.lr_scheduler.MultiStepLR(topt, [15, 80, 150])
    for _ in range(200):
        topt.step()
        ts.step()
        s.step()
        assert abs(o.lr - ts.get_last_lr()[0]) < 1e-15
    assert abs(o.lr - 5e-6) < 1e-12


def test_script_entry_points_exist_with_reference_names_and_defaults():
    """train_ddpm_* / load_test_* carry the reference scripts' names and constants (MSR.py:187-214, NU.py:213-242,
    CO.py:203-232); without a GPU they fail loudly instead of falling back to a CPU path."""
    import inspect
    for mod, train, test, lr, ms in ((D.msr, "train_ddpm_msr", "load_test_msr", 0.005, (100, 150)),
                                     (D.nu, "train_ddpm_nu", "load_test_nu", 0.004, (80, 200)),
                                     (D.co, "train_ddpm_co", "load_test_co", 0.005, (15, 80, 150))):
        sig = inspect.signature(getattr(mod, train)).parameters
        assert sig["epochs"].default == 200 and sig["lr"].default == lr and tuple(sig["milestones"].default) == ms
        assert inspect.signature(getattr(mod, test)).parameters["omega"].default == 500
    if not torch.cuda.is_available():
        from diffsg_b200 import scripts
        with pytest.raises(_lib.DiffsgError):
            scripts.default_device()


# ------------------------------------------------------------------------------ batched baselines
def test_baseline_mirrors_strict_load_reference_checkpoints():
    from baseline_cases import CASES, build_case, golden
    from diffsg_b200 import _lib
    z = golden()
    for name in CASES:
        m = build_case(name, z)                      # load_state_dict(strict=True) inside
        with pytest.raises(_lib.DiffsgError):        # no CPU path
            m(torch.from_numpy(z[f"{name}.x"]))


def test_baseline_mlp_plan_codes():
    from diffsg_b200 import baselines as B
    lin, acts, head = B._mlp_plan(list(B.mtfnn_msr_model(3, 3)))
    assert [m.out_features for m in lin] == [8, 16, 8, 3] and acts == [1, 1, 1, 0] and head == 1
    lin, acts, head = B._mlp_plan(list(B.mtfnn_co_model(9, 3)))
    assert acts == [1, 1, 1, 3] and head == 0
    lin, acts, head = B._mlp_plan(list(B.PPOAgent(6, 5).actor))
    assert acts == [2, 2, 2, 0] and head == 0


def test_cat_params_backward_adds_slices_into_the_parameter_grads():
    """diffsg_b200.train._CatParams (the weight of the shared time table): forward = torch.cat, backward adds every
    slice of the incoming gradient straight into the leaf's .grad (allocating it when absent) and returns nothing."""
    from diffsg_b200.train import _CatParams, _TableGrad
    ps = [torch.nn.Parameter(torch.randn(n, 4)) for n in (3, 1, 5)]
    ps[1].grad = torch.ones_like(ps[1])
    cat = _CatParams.apply(*ps)
    assert torch.equal(cat, torch.cat([p.detach() for p in ps], 0))
    g = torch.arange(cat.numel(), dtype=torch.float32).reshape(cat.shape)
    cat.backward(g)
    assert torch.equal(ps[0].grad, g[:3]) and torch.equal(ps[1].grad, g[3:4] + 1.0) and torch.equal(ps[2].grad, g[4:])
    holder = _TableGrad()
    assert holder.claim() and not holder.claim()            # only the first consumer hands the table gradient back
    buf = holder.buffer(2, 3, "cpu")
    assert buf is holder.buffer(2, 3, "cpu") and float(buf.abs().sum()) == 0.0


def test_training_kernel_index_arithmetic_restated():
    """Host restatement of three pieces of index arithmetic in csrc/train_tc.cu that the GPU tests only exercise on a few
    shapes: (1) FastDiv — q = (x * ceil(2^20 / d)) >> 20 equals x // d for every item counter the staging loops form
    (x < 4096, d <= 256) and the product fits 32 bits; (2) op_off — the UMMA K-major no-swizzle core-matrix layout
    [mn / 8][k / 8][mn % 8][k % 8] is a bijection of a 128 x 64 chunk onto 16 KB of bf16; (3) colsum16 — the
    recursive-halving shuffle butterfly leaves the sum of column 8 b4 + 4 b3 + 2 b2 + b1 in every lane."""
    x = np.arange(4096, dtype=np.uint64)
    for d in range(1, 257):
        m = ((1 << 20) + d - 1) // d
        assert int(x[-1]) * m < 1 << 32
        assert np.array_equal((x * np.uint64(m)) >> np.uint64(20), x // np.uint64(d)), d
    mn, k = np.meshgrid(np.arange(128), np.arange(64), indexing="ij")
    off = (mn >> 3) * (64 * 16) + (k >> 3) * 128 + (mn & 7) * 16 + (k & 7) * 2
    assert off.min() == 0 and off.max() == 128 * 64 * 2 - 2 and len(np.unique(off)) == 128 * 64 and np.all(off % 2 == 0)
    rng = np.random.default_rng(0)
    v = rng.standard_normal((32, 16))                      # [lane][value]
    lanes = np.arange(32)

    def halve(vals, bit, xor):                             # one butterfly step: keep one half, receive the partner's other half
        n = vals.shape[1] // 2
        up = (lanes & bit) != 0
        send = np.where(up[:, None], vals[:, :n], vals[:, n:])
        keep = np.where(up[:, None], vals[:, n:], vals[:, :n])
        return keep + send[lanes ^ xor]

    a = halve(v, 16, 16)
    b = halve(a, 8, 8)
    c = halve(b, 4, 4)
    d1 = halve(c, 2, 2)[:, 0]
    d1 = d1 + d1[lanes ^ 1]
    col = ((lanes >> 4) & 1) * 8 + ((lanes >> 3) & 1) * 4 + ((lanes >> 2) & 1) * 2 + ((lanes >> 1) & 1)
    assert np.allclose(d1, v.sum(0)[col])
