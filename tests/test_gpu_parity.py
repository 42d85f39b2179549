"""GPU parity tests: the CUDA path (through the C-ABI) against the oracle / reference goldens.

Tolerances (BASELINE.json north_star): per-pass eps rel-L2 <= 1e-3 teacher-forced; final
objective within 0.5 % under identical injected noise.  The exact-fp32 engine is held to a
much tighter 2e-5 (fp32 re-association only)."""
import numpy as np
import pytest
import torch

import diffsg_b200 as D
from diffsg_b200 import _lib
from diffsg_b200.engine import philox_normal
from oracle import ddpm_oracle as O
from oracle.standin import CONFIGS

from conftest import load_golden, nu_checkpoint_model, rel_l2, standin_model

pytestmark = pytest.mark.gpu
T = 20
DEV = "cuda:0"


def cuda(a):
    return torch.as_tensor(a).to(DEV)


# engine -> (per-pass eps tolerance, tolerance on a well-conditioned final y)
PRECISIONS = {"fp32": (2e-5, 5e-4), "fp16x3": (2e-5, 5e-4), "fp16x2": (1e-3, 5e-3)}
ENGINE_CASES = [(n, p) for n in CONFIGS for p in PRECISIONS]


def with_precision(ddpm, precision):
    ddpm.model.precision = precision
    assert ddpm.model.engine().precision == precision
    return ddpm


@pytest.mark.parametrize("name,precision", ENGINE_CASES)
def test_unet_forward_matches_reference(name, precision):
    g = load_golden(f"standin_{name}.npz")
    ddpm, cfg = standin_model(name, DEV)
    with_precision(ddpm, precision)
    with torch.no_grad():
        eps = ddpm.model(cuda(g["x"]), cuda(g["ts"]) / T, cuda(g["cond"]), cuda(g["mask"]))
    assert eps.shape == g["eps"].shape
    assert rel_l2(eps.cpu(), g["eps"]) < PRECISIONS[precision][0]


def test_auto_precision_picks_tensor_cores_when_supported():
    ddpm, _ = standin_model("msr80c", DEV)
    assert ddpm.model.engine().precision == "fp16x2" and ddpm.model.engine().tc is not None
    ddpm, _ = standin_model("attn", DEV)           # attention lowers to one accumulate stage per block
    assert ddpm.model.engine().precision == "fp16x2" and ddpm.model.engine().tc is not None
    wide = D.UNet1D(input_dim=200, proj_dim=32, cond_dim=4, dims=(16, 8), is_attn=(False, False), n_blocks=1).to(DEV)
    assert wide.engine().precision == "fp32"       # rows wider than a TMEM region: exact-fp32 engine
    wide.precision = "fp16x3"
    with pytest.raises(_lib.DiffsgError):
        wide.engine()


@pytest.mark.parametrize("cfg", [dict(input_dim=4, proj_dim=24, cond_dim=5, dims=(12, 6), is_attn=(False, True), middle_attn=True, n_blocks=1),
                                 dict(input_dim=7, proj_dim=40, cond_dim=3, dims=(40, 20, 10), is_attn=(False,) * 3, middle_attn=False, n_blocks=2),
                                 dict(input_dim=3, proj_dim=100, cond_dim=9, dims=(72, 36), is_attn=(False,) * 2, middle_attn=False, n_blocks=1)],
                         ids=lambda c: "x".join(map(str, (c["proj_dim"],) + tuple(c["dims"]))))
def test_tensor_core_engine_odd_widths(cfg):
    """Widths that are neither powers of two nor multiples of 16 on the tcgen05 engine: forward against the CPU
    oracle, sampler against the exact-fp32 engine (same injected noise)."""
    from oracle.standin import make_state_dict
    model = D.UNet1D(**cfg)
    sd = make_state_dict({k: v.shape for k, v in model.state_dict().items()}, seed=7)
    model.load_state_dict(sd)
    M = cfg["input_dim"]
    ddpm = D.msr.DDPM(T, model, M, 10.0, 1.0 - D.generate_cosine_schedule(T), DEV, (1, M), {}).to(DEV)
    B = 300
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, M, generator=g)
    c = torch.rand(B, cfg["cond_dim"], generator=g)
    ts = torch.randint(0, T, (1, B), generator=g)
    m = (torch.rand(B, 1, generator=g) > 0.3).float()
    want = O.unet_forward({"model." + k: v for k, v in sd.items()}, x, ts / T, c, m)
    y_T, steps = O.draw_noise(B, (1, M), T, 3)
    outs = {}
    for precision in ("fp32", "fp16x3", "fp16x2"):
        with_precision(ddpm, precision)
        assert (ddpm.model.engine().tc is not None) == (precision != "fp32")
        with torch.no_grad():
            eps = ddpm.model(x.to(DEV), ts.to(DEV) / T, c.to(DEV), m.to(DEV))
        assert rel_l2(eps.cpu(), want) < PRECISIONS[precision][0], precision
        outs[precision] = ddpm.sample(c.to(DEV), 3.0, y_init=y_T.reshape(B, M), noise=torch.stack(steps).reshape(T - 2, B, M)).cpu()
    assert rel_l2(outs["fp16x3"], outs["fp32"]) < 5e-4
    assert rel_l2(outs["fp16x2"], outs["fp32"]) < 3e-2


@pytest.mark.parametrize("precision", list(PRECISIONS))
@pytest.mark.parametrize("B", [1, 7, 8, 9, 127, 129, 200])
def test_unet_forward_ragged_batches(B, precision):
    """Row-group / tile tails (B not a multiple of 8 or 128) and per-row time indices + masks."""
    g = load_golden("standin_nu_like.npz")
    ddpm, cfg = standin_model("nu_like", DEV)
    with_precision(ddpm, precision)
    idx = np.arange(B) % g["x"].shape[0]
    x, cond, ts, mask = g["x"][idx], g["cond"][idx], g["ts"][:, idx], g["mask"][idx]
    with torch.no_grad():
        eps = ddpm.model(cuda(x), cuda(ts) / T, cuda(cond), cuda(mask))
    assert rel_l2(eps.cpu(), g["eps"][idx]) < PRECISIONS[precision][0]


@pytest.mark.parametrize("precision", list(PRECISIONS))
def test_nu_checkpoint_teacher_forced_eps(precision):
    """eps_0 / eps_1 of ddpm_nu_3u.pt at every step, on the reference trajectory."""
    g = load_golden("nu_trace.npz")
    ddpm = with_precision(nu_checkpoint_model(DEV), precision)
    B = g["cond"].shape[0]
    cond = cuda(g["cond"])
    worst = 0.0
    with torch.no_grad():
        for step in range(T):
            i = T - 1 - step
            t = torch.full((1, B), i, device=DEV) / T
            y = cuda(g["y_in"][step])
            e1 = ddpm.model(y, t, cond, torch.ones(B, 1, device=DEV))
            e0 = ddpm.model(y, t, cond, torch.zeros(B, 1, device=DEV))
            worst = max(worst, rel_l2(e1.cpu(), g["eps_1"][step]), rel_l2(e0.cpu(), g["eps_0"][step]))
    assert worst < 1e-3, worst                         # the gate (BASELINE.json north_star)
    assert worst < PRECISIONS[precision][0], worst     # what this engine is expected to deliver
    print(f"[{precision}] worst per-pass eps rel-L2 over 20 steps: {worst:.2e}")


@pytest.mark.parametrize("name,precision", ENGINE_CASES)
def test_sampler_injected_noise_low_guidance(name, precision):
    """Full reverse process with the reference's own draws injected; omega = 3 keeps the
    trajectory well conditioned so the final y itself can be compared."""
    g = load_golden(f"standin_{name}.npz")
    ddpm, cfg = standin_model(name, DEV)
    with_precision(ddpm, precision)
    y0 = ddpm.sample(cuda(g["cond"]), 3.0, y_init=cuda(g["y_T"]), noise=cuda(g["noise"]))
    assert rel_l2(y0.cpu(), g["y0_omega3"]) < PRECISIONS[precision][1]


@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
def test_sampler_records_match_reference(precision):
    g = load_golden("standin_nu_like.npz")
    ddpm, cfg = standin_model("nu_like", DEV)
    with_precision(ddpm, precision)
    B, M = g["cond"].shape[0], cfg["input_dim"]
    rec_y = torch.empty(T, B, M, device=DEV)
    rec_e = torch.empty(T, B, M, device=DEV)
    y = cuda(g["y_T"]).clone()
    ddpm.model.engine().sample(cuda(g["cond"]), y, ddpm.step_coefficients(), T, 3.0, noise=cuda(g["noise"]),
                               rec_y=rec_y, rec_eps=rec_e)
    assert rel_l2(rec_y.cpu(), g["rec_y"]) < 5e-4
    assert rel_l2(rec_e.cpu(), g["rec_eps"]) < 5e-4
    assert rel_l2(y.cpu(), g["y0_omega3"]) < 5e-4


@pytest.mark.parametrize("precision", list(PRECISIONS))
def test_nu_sampler_omega_sweep_reference_stream(precision):
    """ddpm_nu_3u.pt, noise drawn from torch's CPU generator under the same seed
    as the reference run.  Low guidance: compare y0; omega = 500 is chaotic so it is
    judged on the objective below."""
    g = load_golden("nu_trace.npz")
    ddpm = with_precision(nu_checkpoint_model(DEV), precision)
    cond = cuda(g["cond"])
    loose = 10.0 if precision == "fp16x2" else 1.0
    for om, tol in ((0, 2e-4), (1, 2e-4), (10, 1e-3)):
        torch.manual_seed(123)
        y0 = ddpm.sample(cond, float(om))
        assert rel_l2(y0.cpu(), g[f"y0_omega{om}"]) < tol * loose, om


def _nu_eval(ddpm, X, Y, P_sum, seed):
    ddpm.P_sum = P_sum
    ddpm.custom_config = {"width": 400, "height": 400, "P_sum": P_sum}
    torch.manual_seed(seed)
    return D.nu.evaluate(ddpm, X, Y, ddpm.custom_config, omega=500, batch_size=512)


@pytest.mark.parametrize("precision", list(PRECISIONS))
def test_nu_objective_parity_test_split_and_ood(precision):
    """BASELINE config 3: NU 18 mW test split + 30 mW OOD, omega 500, bs 512, same CPU noise
    stream as the reference run: sum-rate ratio within 0.5 % of the reference's."""
    ref = load_golden("nu_objective.npz")
    d = load_golden("nu_data.npz")
    ddpm = with_precision(nu_checkpoint_model(DEV), precision)
    out = _nu_eval(ddpm, d["X_test"], d["Y_test"], 18.0, 123)
    print(f"[{precision}] NU less ratio {out['less_ratio']:.5f} (reference {float(ref['less_ratio_test']):.5f})")
    assert abs(out["less_ratio"] / float(ref["less_ratio_test"]) - 1) < 5e-3
    assert rel_l2(out["true_rate"].cpu(), ref["true_rate_test"]) < 1e-5
    assert abs(float(out["pred_rate"].mean()) / float(ref["pred_rate_test"].mean()) - 1) < 5e-3
    out = _nu_eval(ddpm, d["X_ood"], d["Y_ood"], 30.0, 123)
    assert abs(out["less_ratio"] / float(ref["less_ratio_ood"]) - 1) < 5e-3


def test_nu_record_denoise_path():
    g = load_golden("nu_record.npz")
    ddpm = with_precision(nu_checkpoint_model(DEV), "fp16x3")
    ddpm.record_denoise_path = True
    torch.manual_seed(7)
    y0 = ddpm.sample(cuda(g["cond"]), 500)
    assert ddpm.y_i_record.shape == g["y_i_record"].shape == (32, T * 5)
    assert ddpm.eps_i_record.shape == g["eps_i_record"].shape
    # the first steps are still well conditioned at omega=500: compare them tightly
    k = 4 * 5
    assert rel_l2(ddpm.eps_i_record[:, :k], g["eps_i_record"][:, :k]) < 1e-3
    assert rel_l2(ddpm.y_i_record[:, :k], g["y_i_record"][:, :k]) < 1e-3
    assert y0.shape == g["y0"].shape


def test_philox_stream_matches_oracle_and_sampler_uses_it():
    z = philox_normal(1000, 5, 7, seed=42, offset=3, device=DEV)
    assert rel_l2(z.cpu(), O.philox_normal(1000, 5, 7, seed=42, offset=3)) < 1e-5
    z80 = philox_normal(4096, 80, 3, seed=1, offset=0, device=DEV)
    assert abs(float(z80.mean())) < 0.01 and abs(float(z80.std()) - 1) < 0.01
    # philox mode == injecting the same stream
    ddpm, cfg = standin_model("nu_like", DEV)
    B, M = 64, cfg["input_dim"]
    cond = torch.rand(B, cfg["cond_dim"], device=DEV)
    ddpm.noise_mode, ddpm.philox_seed, ddpm.philox_offset = "philox", 9, 100
    y_a = ddpm.sample(cond, 3.0)
    assert ddpm.philox_offset == 100 + B
    y_T = philox_normal(B, M, T, 9, 100, DEV)
    noise = torch.stack([philox_normal(B, M, i, 9, 100, DEV) for i in range(T - 1, 1, -1)])
    y_b = ddpm.sample(cond, 3.0, y_init=y_T, noise=noise)
    assert rel_l2(y_a.cpu(), y_b.cpu()) < 1e-5


@pytest.mark.parametrize("precision", list(PRECISIONS))
def test_sampler_is_independent_of_sharding_in_phase_b(precision):
    """Rows are independent once the 4 batch-normalised steps are over: sampling two halves
    with the statistics disabled equals sampling the whole (the multi-GPU sharding argument)."""
    ddpm, cfg = standin_model("nu_like", DEV)
    with_precision(ddpm, precision)
    B, M = 300, cfg["input_dim"]
    g = torch.Generator().manual_seed(0)
    cond = torch.rand(B, cfg["cond_dim"], generator=g).to(DEV)
    y_T = torch.randn(B, M, generator=g).to(DEV)
    noise = torch.randn(T - 2, B, M, generator=g).to(DEV)
    eng, coef = ddpm.model.engine(), ddpm.step_coefficients()
    whole = eng.sample(cond, y_T.clone(), coef, T, 3.0, noise=noise, norm_steps=0)
    a = eng.sample(cond[:140].contiguous(), y_T[:140].clone(), coef, T, 3.0, noise=noise[:, :140].contiguous(), norm_steps=0)
    b = eng.sample(cond[140:].contiguous(), y_T[140:].clone(), coef, T, 3.0, noise=noise[:, 140:].contiguous(), norm_steps=0)
    assert torch.equal(whole, torch.cat((a, b)))


def test_ema_fused_update_matches_reference():
    g = load_golden("ema.npz")
    model = D.UNet1D(input_dim=3, proj_dim=16, cond_dim=4, dims=(8, 4, 2), n_blocks=1)
    model.load_state_dict({k[3:]: torch.tensor(v) for k, v in g.items() if k.startswith("p0.")})
    model.to(DEV)
    ema = D.ExponentialMovingAverage(model, 0.9, device=DEV)
    _lib.launch_count(reset=True)
    ema.update_parameters(model)
    assert _lib.launch_count() == 1 and int(ema.n_averaged) == 1
    for k, v in ema.module.state_dict().items():
        assert torch.equal(v.cpu(), torch.tensor(g["p0." + k])), k
    model.load_state_dict({k[3:]: torch.tensor(v) for k, v in g.items() if k.startswith("p1.")})
    ema.update_parameters(model)
    assert int(ema.n_averaged) == 2
    for k, v in ema.module.state_dict().items():
        assert torch.allclose(v.cpu(), torch.tensor(g["avg." + k]), rtol=1e-6, atol=1e-7), k


def test_objective_kernels_match_reference():
    m = load_golden("msr_data.npz")
    lo, hi, W = (float(v) for v in m["scaler"])
    gains = cuda(m["X_test"]) * (hi - lo) + lo
    rate, p = D.objectives.msr_decode_rate(cuda(m["y_rand"]), gains, W, return_alloc=True)
    assert rel_l2(p.cpu() / W, m["dec_rand"]) < 1e-6
    assert rel_l2(rate.cpu(), m["pred_rate_rand"]) < 1e-6
    assert rel_l2(D.objectives.msr_rate(cuda(m["Y_test"]), gains).cpu(), m["true_rate"]) < 1e-6
    assert rel_l2(D.msr.custom_decoder(cuda(m["y_rand"])).cpu(), m["dec_rand"]) < 1e-6
    c = load_golden("co_data.npz")
    lo, hi = (float(v) for v in c["scaler"])
    Xs = cuda(c["X_test"]) * (hi - lo) + lo
    dec = D.co.customized_real_decoder(cuda(c["y_rand"]))
    assert rel_l2(dec.cpu(), c["dec_rand"]) < 1e-6 and float(dec[:5].abs().sum()) == 0.0
    assert rel_l2(D.co.cost_calc(Xs, dec).cpu(), c["pred_cost_rand"]) < 1e-5
    assert rel_l2(D.co.cost_calc(Xs, cuda(c["Y_test"])).cpu(), c["true_cost"]) < 1e-5
    n, d = load_golden("nu_objective.npz"), load_golden("nu_data.npz")
    dec = D.nu.custom_decoder(cuda(n["y0_test"]), 400, 400, 18.0)
    assert rel_l2(dec.cpu(), O.nu_decode(torch.tensor(n["y0_test"]), 400, 400, 18.0)) < 1e-6
    rate = D.nu.rate_calc(dec, cuda(d["X_test"]) * 400.0)
    assert rel_l2(rate.cpu(), n["pred_rate_test"]) < 1e-5
    # M = 80 (sub-warp groups of 32 lanes, strided columns)
    g = torch.Generator().manual_seed(1)
    y80, g80 = torch.randn(513, 80, generator=g), torch.rand(513, 80, generator=g) * 2 + 0.5
    want = O.msr_rate(20.0 * O.msr_decode(y80), g80)
    assert rel_l2(D.objectives.msr_decode_rate(y80.to(DEV), g80.to(DEV), 20.0).cpu(), want) < 1e-6


@pytest.mark.parametrize("name", ["attn", "nu_like"])
def test_repack_after_parameter_update(name):
    ddpm, cfg = standin_model(name, DEV)
    g = load_golden(f"standin_{name}.npz")
    args = (cuda(g["x"]), cuda(g["ts"]) / T, cuda(g["cond"]), cuda(g["mask"]))
    with torch.no_grad():
        a = ddpm.model(*args)
        ddpm.model.final.bias.add_(1.0)
        b = ddpm.model(*args)
    assert torch.allclose(b, a + 1.0, atol=1e-4)


def test_errors_are_loud():
    ddpm, cfg = standin_model("attn", DEV)
    with pytest.raises(ValueError):
        ddpm.sample(torch.rand(8, cfg["cond_dim"], device=DEV), 1.0, y_init=torch.zeros(8, cfg["input_dim"]),
                    noise=torch.zeros(3, 8, cfg["input_dim"]))
    with pytest.raises(_lib.DiffsgError):
        D.objectives.nu_rate(torch.rand(4, 40, device=DEV), torch.rand(4, 76, device=DEV))  # K > 32


def test_training_loss_and_grads_match_reference_autograd():
    """eps-MSE loss and every parameter gradient for a fixed (ts, noise, mask) triple
    against the reference's autograd (golden from oracle/make_golden.py)."""
    g = load_golden("train_step.npz")
    ddpm, cfg = standin_model("nu_like", DEV)
    y, cond, noise, mask = (cuda(g[k]) for k in ("y", "cond", "noise", "mask"))
    ts = cuda(g["ts"])
    y_t = torch.squeeze(ddpm.sqrt_alphas_cumprod[ts, None] * y + ddpm.sqrt_one_minus_alphas_cumprod[ts, None] * noise)
    ddpm.zero_grad()
    _lib.launch_count(reset=True)
    loss = ddpm.loss_from(y_t, ts, cond, mask, noise)
    loss.backward()
    assert _lib.launch_count() >= 2 * 67          # every Linear ran the tcgen05 forward node and the one-launch backward node
    assert abs(float(loss.detach()) / float(g["loss"]) - 1) < 1e-5
    worst, who, errs = 0.0, None, []
    for name, p in ddpm.model.named_parameters():
        want = g["grad." + name]
        if np.abs(want).max() == 0:
            assert float(p.grad.abs().max()) < 1e-12, name
            continue
        e = rel_l2(p.grad.cpu(), want)
        errs.append(e)
        if e > worst:
            worst, who = e, name
    print(f"worst parameter-gradient rel-L2 {worst:.2e} ({who}), median {float(np.median(errs)):.2e}")
    # operands are bf16 hi+lo (16 significant bits, csrc/train_tc.cu): well inside the 1e-3 contraction gate of
    # BASELINE.json's north star (it prescribes bf16 / tf32 contractions); fp32 cuBLAS measured ~2e-5 here
    assert worst < 5e-4, (worst, who)
    assert float(np.median(errs)) < 1e-4


def test_training_forward_equals_inference_engine():
    ddpm, cfg = standin_model("co", DEV)
    g = load_golden("standin_co.npz")
    args = (cuda(g["x"]), cuda(g["ts"]) / T, cuda(g["cond"]), cuda(g["mask"]))
    eps_train = ddpm.model(*args)                      # grad mode: training graph
    assert eps_train.requires_grad
    assert rel_l2(eps_train.detach().cpu(), g["eps"]) < 2e-5


def test_data_parallel_trainer_single_gpu_step_reduces_loss():
    from diffsg_b200.parallel import DataParallelTrainer
    d = load_golden("nu_data.npz")
    ddpm, cfg = standin_model("nu_like", DEV)
    ddpm.apply(D.init_weights)
    tr = DataParallelTrainer(ddpm, lr=1e-3)
    X, Y = cuda(d["X_train_head"]), cuda(d["Y_train_head"])
    torch.manual_seed(0)
    first = [float(tr.step(Y[i:i + 512], X[i:i + 512])) for i in (0, 512)]
    for _ in range(40):
        for i in (0, 512):
            last = float(tr.step(Y[i:i + 512], X[i:i + 512]))
    assert last < 0.8 * first[0], (first, last)
    # parameters are still views of the flat buffer and the sampler sees the updated weights
    assert all(p.data_ptr() >= tr.flat.flat.data_ptr() for p in ddpm.model.parameters())
    y0 = ddpm.sample(X[:64], 1.0)
    assert torch.isfinite(y0).all()


@pytest.mark.parametrize("precision", ["fp32", "fp16x2"])
def test_forward_steps_inference_matches_forward(precision):
    """UNet1D.forward_steps without grad = the engine forward on the cached step table (no torch.unique sync):
    bit-identical to forward(x, ts / T, ...)."""
    ddpm, cfg = standin_model("nu_like")
    ddpm = ddpm.to(DEV)
    ddpm.model.precision = precision
    B = 300
    g = torch.Generator().manual_seed(4)
    x = torch.randn(B, cfg["input_dim"], generator=g).to(DEV)
    c = torch.rand(B, cfg["cond_dim"], generator=g).to(DEV)
    ts = torch.randint(0, T, (1, B), generator=g).to(DEV)
    m = torch.ones(B, 1, device=DEV)
    with torch.no_grad():
        a = ddpm.model.forward_steps(x, ts, T, c, m)
        b = ddpm.model(x, ts / T, c, m)
    assert torch.equal(a, b)


def test_training_time_path_hoisting_is_exact():
    """UNet1D.forward_steps (time path evaluated on the T grid rows, gathered per sample) against the plain
    training graph (per-sample time path, the reference's formulation): same eps, same parameter gradients."""
    from diffsg_b200.train import unet_forward_train
    ddpm, cfg = standin_model("nu_like")
    ddpm = ddpm.to(DEV)
    model = ddpm.model
    B = 2048
    g = torch.Generator().manual_seed(9)
    x = torch.randn(B, cfg["input_dim"], generator=g).to(DEV)
    c = torch.rand(B, cfg["cond_dim"], generator=g).to(DEV)
    ts = torch.randint(0, T, (1, B), generator=g).to(DEV)
    m = (torch.rand(B, 1, generator=g) > 0.1).float().to(DEV)
    w = torch.randn(B, cfg["input_dim"], generator=g).to(DEV)
    grads = []
    for hoist in (False, True):
        model.zero_grad(set_to_none=True)
        if hoist:
            eps = model.forward_steps(x, ts, T, c, m)
        else:
            eps = unet_forward_train(model, x, ts / T, c, m)
        (eps * w).sum().backward()
        grads.append((eps.detach().clone(), {k: p.grad.detach().clone() for k, p in model.named_parameters()}))
    assert rel_l2(grads[1][0].cpu(), grads[0][0].cpu()) < 1e-5
    for k in grads[0][1]:
        assert rel_l2(grads[1][1][k].cpu(), grads[0][1][k].cpu()) < 2e-4, k


def test_training_long_schedule_gathers_the_time_rows_once():
    """More time rows than the wgrad kernel's one-hot columns (T = 40 > 31): the hoisted time path is gathered once with
    index_select and added row by row; same eps and gradients as the per-sample time path."""
    from diffsg_b200.train import MAX_GATHER_ROWS, unet_forward_train
    T_long = 40
    assert T_long > MAX_GATHER_ROWS
    ddpm, cfg = standin_model("nu_like")
    model = ddpm.to(DEV).model
    B = 333
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, cfg["input_dim"], generator=g).to(DEV)
    c = torch.rand(B, cfg["cond_dim"], generator=g).to(DEV)
    ts = torch.randint(0, T_long, (1, B), generator=g).to(DEV)
    m = (torch.rand(B, 1, generator=g) > 0.1).float().to(DEV)
    w = torch.randn(B, cfg["input_dim"], generator=g).to(DEV)
    got = []
    for hoist in (False, True):
        model.zero_grad(set_to_none=True)
        eps = model.forward_steps(x, ts, T_long, c, m) if hoist else unet_forward_train(model, x, ts / T_long, c, m)
        (eps * w).sum().backward()
        got.append((eps.detach().clone(), {k: p.grad.detach().clone() for k, p in model.named_parameters()}))
    assert rel_l2(got[1][0].cpu(), got[0][0].cpu()) < 1e-5
    for k in got[0][1]:
        assert rel_l2(got[1][1][k].cpu(), got[0][1][k].cpu()) < 3e-4, k


def test_trainer_cuda_graph_mode():
    """cuda_graph=True: capturing must not change the model (lr-0 warm-up, optimiser state reset), and the
    replayed step trains like the eager one (different RNG streams: compare the loss level, not bits)."""
    from diffsg_b200.parallel import DataParallelTrainer
    kind, cfg = CONFIGS["nu_like"]
    M = cfg["input_dim"]
    g = torch.Generator().manual_seed(3)
    X = torch.rand(512, cfg["cond_dim"], generator=g).to(DEV)
    Y = torch.rand(512, M, generator=g).to(DEV)
    finals = {}
    for mode in (False, True):
        torch.manual_seed(0)
        model = D.UNet1D(**cfg)
        ddpm = D.msr.DDPM(T, model, M, 10.0, 1.0 - D.generate_cosine_schedule(T), DEV, (1, M), {}).to(DEV)
        ddpm.apply(D.init_weights)
        tr = DataParallelTrainer(ddpm, lr=1e-3, cuda_graph=mode)
        if mode:
            before = tr.flat.flat.clone()
            tr._graphs[((512, M), (512, cfg["cond_dim"]), False)] = tr._capture(Y, X)
            assert torch.equal(before, tr.flat.flat)
            assert float(tr.opt.exp_avg.abs().sum()) == 0 and float(tr.opt.exp_avg_sq.abs().sum()) == 0 and int(tr.opt.step_dev) == 0
        losses = [tr.step(Y, X) for _ in range(150)]
        first, last = float(torch.stack(losses[:10]).mean()), float(torch.stack(losses[-30:]).mean())
        assert last < 0.7 * first, (mode, first, last)
        assert torch.isfinite(tr.flat.flat).all()
        finals[mode] = last
        y0 = ddpm.sample(X[:64], 1.0)                 # the inference engine sees the graph-updated weights
        assert torch.isfinite(y0).all()
    assert abs(finals[True] / finals[False] - 1) < 0.4, finals      # different RNG streams: same loss LEVEL


def _train_standin(kind, X, Y, cfg_net, steps, lr=1e-3):
    """Stand-in checkpoint for a configuration whose real checkpoint is missing from the reference
    repo: the reference recipe (init_weights, Adam, bs 512) at lr 1e-3, trained here
    with the data-parallel trainer."""
    from diffsg_b200.parallel import DataParallelTrainer
    torch.manual_seed(0)
    model = D.UNet1D(**cfg_net)
    alphas = 1.0 - D.generate_cosine_schedule(T)
    M = cfg_net["input_dim"]
    if kind == "co":
        ddpm = D.co.DDPM(T, model, M, alphas, DEV, (1, M), {}, 0.1, 0.9999, 10, 5, False)
    else:
        ddpm = D.msr.DDPM(T, model, M, 10.0, alphas, DEV, (1, M), {}, 0.1, 0.9999, 10, 5, False)
    ddpm.apply(D.init_weights)
    ddpm.to(DEV)
    tr = DataParallelTrainer(ddpm, lr=lr)
    n = X.shape[0]
    g = torch.Generator().manual_seed(1)
    losses = []
    for s in range(steps):
        idx = torch.randint(0, n, (512,), generator=g).to(DEV)
        losses.append(float(tr.step(Y[idx], X[idx])))
    # single-batch losses are noisy (random t, noise, mask): compare window means
    return ddpm, sum(losses[:10]) / 10, sum(losses[-50:]) / 50


def _oracle_state(ddpm):
    return {k: v.detach().cpu().clone() for k, v in ddpm.state_dict().items() if not k.startswith("ema.")}


@pytest.mark.parametrize("kind", ["msr", "co"])
def test_objective_parity_on_trained_standin(kind):
    """BASELINE configs 1 / 4: objective of the sampled solutions on the bundled test split, same weights
    and identical injected noise, CUDA engines vs the CPU oracle: within 0.5 % (stand-in checkpoint)."""
    if kind == "msr":
        d = load_golden("msr_data.npz")
        net = CONFIGS["msr3c"][1]
        lo, hi, W = (float(v) for v in d["scaler"])
    else:
        d = load_golden("co_data.npz")
        net = CONFIGS["co"][1]
        lo, hi = (float(v) for v in d["scaler"])
    Xtr, Ytr = cuda(d["X_train"]), cuda(d["Y_train"])
    ddpm, first, last = _train_standin(kind, Xtr, Ytr, net, steps=400)
    assert last < 0.7 * first, (first, last)           # it actually learned something
    B, M = min(len(d["X_test"]), 1024), net["input_dim"]
    Xte = torch.tensor(d["X_test"][:B])
    sd = _oracle_state(ddpm)
    omegas = (500.0,) if kind == "msr" else (0.0, 10.0, 500.0)      # CO: guidance-weight sweep (config 4)
    for omega in omegas:
        y_T, steps = O.draw_noise(B, (1, M), T, 11)
        with torch.no_grad():
            y_ref = O.sample(sd, T, Xte, omega, y_T, steps)
        raw = Xte * (hi - lo) + lo
        if kind == "msr":
            obj_ref = O.msr_rate(W * O.msr_decode(y_ref), raw)
        else:
            obj_ref = O.co_cost(raw, O.co_decode(y_ref))
        for precision in ("fp32", "fp16x3", "fp16x2"):
            ddpm.model.precision = precision
            y0 = ddpm.sample(Xte.to(DEV), omega, y_init=y_T.reshape(B, M), noise=torch.stack(steps).reshape(T - 2, B, M))
            if kind == "msr":
                obj = D.objectives.msr_decode_rate(y0, raw.to(DEV), W)
            else:
                obj = D.co.cost_calc(raw.to(DEV), D.co.customized_real_decoder(y0))
            obj = obj.cpu().reshape(-1)
            ref = obj_ref.reshape(-1)
            ratio = float(obj.mean()) / float(ref.mean())
            print(f"[{kind} omega={omega} {precision}] objective mean {float(obj.mean()):.5f} vs oracle {float(ref.mean()):.5f}")
            if kind == "co":
                # cost_calc thresholds the allocation at 0.1 (CO:261): a row whose decoded allocation sits on the
                # threshold flips its offloading decision under ANY rounding difference (the fp32 engine does too) and
                # its cost jumps.  Parity = same decisions on all but a handful of rows, and the mean within 0.5 % there.
                same = ((D.co.customized_real_decoder(y0).cpu() > 0.1) == (O.co_decode(y_ref) > 0.1)).all(dim=1)
                flipped = 1.0 - float(same.float().mean())
                print(f"    decision pattern differs on {flipped * 100:.2f} % of rows")
                assert flipped < 0.02, (kind, omega, precision, flipped)
                ratio = float(obj[same].mean()) / float(ref[same].mean())
                assert abs(float(obj.mean()) / float(ref.mean()) - 1) < 2e-2, (kind, omega, precision)
            assert abs(ratio - 1) < 5e-3, (kind, omega, precision, ratio)


@pytest.mark.parametrize("name,precision", [("attn", "fp32"), ("attn", "fp16x3"), ("nu_like", "fp16x3"), ("nu_like", "fp32")])
def test_short_schedule_every_step_renormalised(name, precision):
    """T = 3 < 5: the reference's `i > T - 5` makes EVERY step (including the last) re-normalise over the
    batch, and no step draws noise except i = 2 (`i > 1`)."""
    from oracle.standin import make_state_dict
    Ts = 3
    kind, cfg = CONFIGS[name]
    model = D.UNet1D(**cfg)
    sdm = make_state_dict({k: v.shape for k, v in model.state_dict().items()}, seed=1234)
    model.load_state_dict(sdm)
    alphas = 1.0 - D.generate_cosine_schedule(Ts)
    M = cfg["input_dim"]
    ddpm = D.msr.DDPM(Ts, model, M, 10.0, alphas, DEV, (1, M), {}).to(DEV)
    ddpm.model.precision = precision
    full = {"model." + k: v for k, v in sdm.items()}
    full.update(O.ddpm_buffers(alphas))
    B = 37
    g = torch.Generator().manual_seed(4)
    cond = torch.rand(B, cfg["cond_dim"], generator=g)
    y_T, steps = O.draw_noise(B, (1, M), Ts, 9)
    assert len(steps) == 1
    with torch.no_grad():
        want = O.sample(full, Ts, cond, 2.0, y_T, steps)
    y0 = ddpm.sample(cond.to(DEV), 2.0, y_init=y_T.reshape(B, M), noise=torch.stack(steps).reshape(1, B, M))
    assert rel_l2(y0.cpu(), want) < 5e-4
    # the output of a batch-normalised last step has zero mean / unit (unbiased) variance over the batch
    assert abs(float(y0.mean())) < 1e-4 and abs(float(y0.var()) - 1.0) < 1e-3


@pytest.mark.parametrize("precision", ["fp32", "fp16x2"])
def test_sharded_sampling_with_whole_batch_statistics(precision):
    """`sample(..., stats_group=...)`: two row shards of ONE batch (two plans driven from two threads, the
    all-reduce emulated with a barrier) reproduce the un-sharded call, whose four re-normalised steps use the
    statistics of the whole batch (reference MSR.py:136-137); without it the shards would normalise separately."""
    import copy
    import threading
    ddpm, cfg = standin_model("nu_like", DEV)
    with_precision(ddpm, precision)
    M, Cd = cfg["input_dim"], cfg["cond_dim"]
    B, cut = 700, 300                                   # ragged shards
    g = torch.Generator().manual_seed(21)
    cond = torch.rand(B, Cd, generator=g)
    y_T = torch.randn(B, M, generator=g)
    noise = torch.randn(T - 2, B, M, generator=g)
    full = ddpm.sample(cond.to(DEV), 3.0, y_init=y_T, noise=noise).cpu()
    sep = torch.cat([ddpm.sample(cond[s].to(DEV), 3.0, y_init=y_T[s], noise=noise[:, s]).cpu()
                     for s in (slice(0, cut), slice(cut, B))])
    assert rel_l2(sep, full) > 1e-3                      # per-shard statistics really are a different computation

    shards = [slice(0, cut), slice(cut, B)]
    barrier, slots, outs, errs = threading.Barrier(2), [None, None], [None, None], []

    def make_reduce(rank):
        def reduce(t):
            torch.cuda.synchronize()
            slots[rank] = t.clone()
            barrier.wait()
            total = slots[0] + slots[1]
            barrier.wait()
            t.copy_(total)
        return reduce

    def work(rank):
        try:
            torch.cuda.set_device(0)
            m = copy.deepcopy(ddpm)
            m.model.precision = precision
            s = shards[rank]
            outs[rank] = m.sample(cond[s].to(DEV), 3.0, y_init=y_T[s], noise=noise[:, s], stats_group=make_reduce(rank)).cpu()
        except Exception as e:                           # pragma: no cover
            errs.append(e)
            barrier.abort()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(2)]
    [t.start() for t in threads]
    [t.join(timeout=120) for t in threads]
    assert not errs, errs
    got = torch.cat(outs)
    assert rel_l2(got, full) < (1e-5 if precision == "fp32" else 1e-4)


def test_degenerate_batches():
    ddpm, cfg = standin_model("nu_like", DEV)
    M, C = cfg["input_dim"], cfg["cond_dim"]
    ddpm.noise_mode = "philox"
    y = ddpm.sample(torch.rand(1, C, device=DEV), 3.0)
    assert y.shape == (M,) and torch.isfinite(y).all()          # torch.squeeze semantics of the reference (MSR.py:116)
    with torch.no_grad():
        eps = ddpm.model(torch.rand(1, M, device=DEV), torch.zeros(1, 1, device=DEV), torch.rand(1, C, device=DEV),
                         torch.ones(1, 1, device=DEV))
    assert eps.shape == (1, M)
    y = ddpm.sample(torch.rand(129, C, device=DEV), 3.0)         # one full tile + one row
    assert y.shape == (129, M) and torch.isfinite(y).all()
    for precision in ("fp16x2", "fp32"):                          # empty batch: empty result, no launch, no error
        ddpm.model.precision = precision
        y = ddpm.sample(torch.rand(0, C, device=DEV), 3.0)
        assert y.numel() == 0
        with torch.no_grad():
            eps = ddpm.model(torch.rand(0, M, device=DEV), torch.zeros(1, 0, device=DEV), torch.rand(0, C, device=DEV),
                             torch.ones(0, 1, device=DEV))
        assert eps.shape == (0, M)
