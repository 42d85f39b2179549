"""Pin the oracle against the UNMODIFIED reference and emit tests/golden/*.

Run in the build container (where /root/reference is mounted read-only):

    python oracle/make_golden.py

It imports the reference's own modules, asserts that `oracle/ddpm_oracle.py` reproduces
them bit-for-bit (UNet forward, full sampler under the same seed, objectives), and writes
small fixtures that travel to the GPU box, where /root/reference does not exist.
TEST INFRASTRUCTURE ONLY — nothing under diffsg_b200/ imports this.
"""
from __future__ import annotations

import io
import json
import os
import sys
from contextlib import redirect_stdout
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
REF = Path(os.environ.get("DIFFSG_REFERENCE", "/root/reference"))
GOLD = ROOT / "tests" / "golden"
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(REF))

from oracle import ddpm_oracle as O  # noqa: E402
from oracle.standin import CONFIGS, make_state_dict  # noqa: E402

with redirect_stdout(io.StringIO()):
    from ddpm_opt.UNetCF import UNet1D  # noqa: E402
    from ddpm_opt.diffusion import generate_cosine_schedule  # noqa: E402
    from ddpm_opt.ema import ExponentialMovingAverage  # noqa: E402
    import ddpm_opt.classifier_free_MSR as RMSR  # noqa: E402
    import ddpm_opt.classifier_free_NU as RNU  # noqa: E402
    import ddpm_opt.classifier_free_CO as RCO  # noqa: E402

torch.set_num_threads(8)
T = 20


def ref_ddpm(kind, model, cfg):
    alphas = 1.0 - generate_cosine_schedule(T)
    M = cfg["input_dim"]
    if kind == "nu":
        return RNU.DDPM(T, model, 3, 18.0, alphas, "cpu", (1, M), {}, 0.1, 0.9999, 10, 5, False)
    if kind == "co":
        return RCO.DDPM(T, model, M, alphas, "cpu", (1, M), {}, 0.1, 0.9999, 10, 5, False)
    return RMSR.DDPM(T, model, M, 10.0, alphas, "cpu", (1, M), {}, 0.1, 0.9999, 10, 5, False)


def save(name, **arrays):
    out = {k: (v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrays.items()}
    np.savez_compressed(GOLD / name, **out)
    print(f"wrote {name}: " + ", ".join(f"{k}{tuple(v.shape)}" for k, v in out.items()))


@torch.no_grad()
def main():
    GOLD.mkdir(parents=True, exist_ok=True)
    manifest = {}

    # ---- (5) schedule -----------------------------------------------------------------
    betas = generate_cosine_schedule(T)
    assert np.array_equal(betas, O.cosine_betas(T))
    bufs = O.ddpm_buffers(1.0 - betas)
    save("schedule_T20.npz", betas64=betas, **bufs)

    # ---- NU: the one real checkpoint ------------------------------------------------------
    sd = torch.load(REF / "ckpts" / "ddpm_nu_3u.pt", map_location="cpu")
    manifest["nu_ckpt"] = {k: list(v.shape) for k, v in sd.items()}
    nu_cfg = dict(input_dim=5, proj_dim=32, cond_dim=6, dims=(32, 16, 8), is_attn=(False,) * 3,
                  middle_attn=False, n_blocks=2)
    model = UNet1D(**nu_cfg)
    ddpm = ref_ddpm("nu", model, nu_cfg)
    ddpm.load_state_dict(sd)
    for k, v in bufs.items():
        assert torch.equal(v, sd[k]), k
    keep = {k: v for k, v in sd.items() if not k.startswith("ema.")}
    save("nu_ckpt.npz", **keep)

    # dataset slices (the GPU box has no /root/reference)
    with redirect_stdout(io.StringIO()):
        Xtr, Ytr, Xte, Yte, Rte, nucfg = RNU.nu_data_load(str(REF / "datasets" / "3u_18mW_10000samples.csv"), 400, 400)
    ood_link = Path("/tmp/3u_30mW_1000samples.csv")
    if not ood_link.exists():
        ood_link.symlink_to(REF / "datasets" / "3u_30mW_1000samples_ood.csv")
    with redirect_stdout(io.StringIO()):
        _, _, Xo, Yo, Ro, oodcfg = RNU.nu_data_load(str(ood_link), 400, 400)
    src = np.array(__import__("pandas").read_csv(ood_link, header=None), dtype=np.float64)
    Xo_all, Yo_all = src[:, :6].copy(), src[:, 6:11].copy()
    Xo_all /= 400.0
    Yo_all[:, :2] /= 400.0
    Yo_all[:, 2:] /= 30.0
    save("nu_data.npz", X_test=Xte.astype(np.float32), Y_test=Yte.astype(np.float32), R_test=Rte.astype(np.float32),
         X_train_head=Xtr[:1024].astype(np.float32), Y_train_head=Ytr[:1024].astype(np.float32),
         X_ood=Xo_all.astype(np.float32), Y_ood=Yo_all.astype(np.float32))

    # (1) teacher-forced eps trace + (2) full sampler, injected noise, seed 123
    B = 256
    cond = torch.tensor(Xte[:B], dtype=torch.float32)
    torch.manual_seed(123)
    y_ref = ddpm.sample(cond, 500)
    y_T, steps = O.draw_noise(B, (1, 5), T, 123)
    trace = {}
    y_or = O.sample(sd, T, cond, 500, y_T, steps, trace=trace)
    assert torch.equal(y_ref, y_or), "oracle sampler != reference sampler (NU, seed 123)"
    finals = {}
    for om in (0.0, 1.0, 10.0, 100.0, 500.0):
        torch.manual_seed(123)
        r = ddpm.sample(cond, om)
        o = O.sample(sd, T, cond, om, y_T, steps)
        assert torch.equal(r, o), om
        finals[f"y0_omega{int(om)}"] = o
    save("nu_trace.npz", cond=cond, y_T=y_T.reshape(B, 5), noise=torch.stack(steps).reshape(T - 2, B, 5),
         y_in=torch.stack(trace["y_in"]), eps_0=torch.stack(trace["eps_0"]), eps_1=torch.stack(trace["eps_1"]),
         **finals)

    # recorded trajectory (record_denoise_path) on a small batch
    ddpm.record_denoise_path = True
    RNU.width, RNU.height = 400, 400  # the reference reads these module globals (NU.py:174)
    torch.manual_seed(7)
    y_rec = ddpm.sample(cond[:32], 500)
    save("nu_record.npz", cond=cond[:32], y0=y_rec, y_i_record=ddpm.y_i_record, eps_i_record=ddpm.eps_i_record)
    ddpm.record_denoise_path = False

    # (3) objective on the full test split and the OOD set, seed 123, bs 512 like load_test_nu
    def nu_eval(X, Y, P_sum, seed):
        Xt = torch.tensor(X, dtype=torch.float32)
        torch.manual_seed(seed)
        Yp = torch.cat([ddpm.sample(Xt[i:i + 512], 500) for i in range(0, Xt.shape[0], 512)])
        # oracle path with the same draws
        torch.manual_seed(seed)
        Yo_ = []
        for i in range(0, Xt.shape[0], 512):
            b = min(512, Xt.shape[0] - i)
            yT = torch.randn(b, 1, 5)
            st = [torch.randn(b, 1, 5) for _ in range(T - 2)]
            Yo_.append(O.sample(sd, T, Xt[i:i + 512], 500, yT, st))
        assert torch.equal(Yp, torch.cat(Yo_))
        Xs = Xt.clone()
        Xs *= 400.0
        Yt = torch.tensor(Y, dtype=torch.float32).clone()
        Yt[:, :2] *= 400.0
        Yt[:, 2:] *= P_sum
        dec_ref = RNU.custom_decoder(Yp, 400, 400, P_sum)
        assert torch.equal(dec_ref, O.nu_decode(Yp, 400, 400, P_sum))
        pr, tr = O.nu_rate(dec_ref, Xs), O.nu_rate(Yt, Xs)
        return Yp, pr, tr

    Yp, pr, tr = nu_eval(Xte, Yte, 18.0, 123)
    # the reference's rate_calc is a Python double loop: check the vectorised oracle on a slice
    Xs = torch.tensor(Xte[:200], dtype=torch.float32) * 400.0
    dec = RNU.custom_decoder(Yp, 400, 400, 18.0)[:200]
    rr = RNU.rate_calc(dec, Xs)
    assert torch.allclose(rr, O.nu_rate(dec, Xs), rtol=1e-6, atol=1e-9), "nu_rate oracle != reference rate_calc"
    Ypo, pro, tro = nu_eval(Xo_all, Yo_all, 30.0, 123)
    save("nu_objective.npz", y0_test=Yp, pred_rate_test=pr, true_rate_test=tr,
         less_ratio_test=float(pr.sum() / tr.sum()), y0_ood=Ypo, pred_rate_ood=pro, true_rate_ood=tro,
         less_ratio_ood=float(pro.sum() / tro.sum()), rate_calc_ref_first200=rr)
    print("NU less ratio test/ood:", float(pr.sum() / tr.sum()), float(pro.sum() / tro.sum()))

    # ---- stand-in configurations (deterministic weights; see oracle/standin.py) -----------
    for name, (kind, cfg) in CONFIGS.items():
        model = UNet1D(**cfg)
        ddpm_r = ref_ddpm(kind, model, cfg)
        manifest[name] = {k: list(v.shape) for k, v in ddpm_r.state_dict().items()}
        sdm = make_state_dict({k: v.shape for k, v in model.state_dict().items()}, seed=1234)
        model.load_state_dict(sdm)
        full = {"model." + k: v for k, v in sdm.items()}
        full.update(bufs)
        Bc = 48
        g = torch.Generator().manual_seed(99)
        x = torch.randn(Bc, cfg["input_dim"], generator=g)
        cond = torch.rand(Bc, cfg["cond_dim"], generator=g)
        ts = torch.randint(0, T, (1, Bc), generator=g)
        mask = (torch.rand(Bc, 1, generator=g) < 0.7).float()
        eps_ref = model(x, ts / T, cond, mask)
        eps_or = O.unet_forward(full, x, ts / T, cond, mask)
        assert torch.equal(eps_ref, eps_or), f"oracle forward != reference forward ({name})"
        torch.manual_seed(5)
        y_ref = ddpm_r.sample(cond, 3.0)
        y_T, steps = O.draw_noise(Bc, (1, cfg["input_dim"]), T, 5)
        trace = {}
        y_or, rec_y, rec_e = O.sample(full, T, cond, 3.0, y_T, steps, record=True, trace=trace)
        assert torch.equal(y_ref.reshape(y_or.shape), y_or), f"oracle sampler != reference ({name})"
        save(f"standin_{name}.npz", x=x, cond=cond, ts=ts, mask=mask, eps=eps_ref,
             y_T=y_T.reshape(Bc, -1), noise=torch.stack(steps).reshape(T - 2, Bc, -1), y0_omega3=y_or,
             rec_y=rec_y, rec_eps=rec_e, eps_0=torch.stack(trace["eps_0"]), eps_1=torch.stack(trace["eps_1"]),
             y_in=torch.stack(trace["y_in"]))

    # ---- objectives: MSR / CO decoders and costs on the bundled data ----------------------
    with redirect_stdout(io.StringIO()):
        Xtr, Ytr, Xte, Yte, mcfg = RMSR.msr_data_load(str(REF / "datasets" / "3c_10w_10000samples.csv"))
    g = torch.Generator().manual_seed(3)
    yp = torch.randn(3000, 3, generator=g) * 2.0
    dec = RMSR.custom_decoder(yp)
    assert torch.equal(dec, O.msr_decode(yp))
    gains = torch.tensor(Xte, dtype=torch.float32) * (mcfg["scaler_max"] - mcfg["scaler_min"]) + mcfg["scaler_min"]
    pred_rate = O.msr_rate(mcfg["W"] * dec, gains)
    true_rate = O.msr_rate(torch.tensor(Yte, dtype=torch.float32), gains)
    save("msr_data.npz", X_test=Xte.astype(np.float32), Y_test=Yte.astype(np.float32),
         X_train=Xtr.astype(np.float32), Y_train=Ytr.astype(np.float32),
         scaler=np.array([mcfg["scaler_min"], mcfg["scaler_max"], mcfg["W"]]), y_rand=yp, dec_rand=dec,
         pred_rate_rand=pred_rate, true_rate=true_rate)
    print("MSR mean true rate:", float(true_rate.mean()))

    with redirect_stdout(io.StringIO()):
        Xtr, Ytr, Xte, Yte, ccfg = RCO.co_data_load(str(REF / "datasets" / "3nodes_2000samples_ood.csv"))
    Xs = torch.tensor(Xte, dtype=torch.float32) * (ccfg["scaler_max"] - ccfg["scaler_min"]) + ccfg["scaler_min"]
    Yt = torch.tensor(Yte, dtype=torch.float32)
    true_cost = RCO.cost_calc(Xs, Yt)
    assert torch.equal(true_cost, O.co_cost(Xs, Yt))
    yp = torch.randn(Xs.shape[0], 3, generator=g) * 4.0
    yp[:5] = -20.0  # exercises the all-below--10 guard
    dec = RCO.customized_real_decoder(yp)
    assert torch.equal(dec, O.co_decode(yp))
    pred_cost = RCO.cost_calc(Xs, dec)
    assert torch.equal(pred_cost, O.co_cost(Xs, dec))
    save("co_data.npz", X_test=Xte.astype(np.float32), Y_test=Yte.astype(np.float32),
         X_train=Xtr.astype(np.float32), Y_train=Ytr.astype(np.float32),
         scaler=np.array([ccfg["scaler_min"], ccfg["scaler_max"]]), y_rand=yp, dec_rand=dec,
         pred_cost_rand=pred_cost, true_cost=true_cost)
    print("CO mean true cost:", float(true_cost.mean()))

    # ---- (6) EMA: two update_parameters calls ---------------------------------------------
    torch.manual_seed(11)
    small = UNet1D(input_dim=3, proj_dim=16, cond_dim=4, dims=(8, 4, 2), n_blocks=1)
    ema = ExponentialMovingAverage(small, 0.9)
    p0 = {k: v.clone() for k, v in small.state_dict().items()}
    ema.update_parameters(small)
    for p in small.parameters():
        p.add_(torch.randn_like(p) * 0.1)
    p1 = {k: v.clone() for k, v in small.state_dict().items()}
    ema.update_parameters(small)
    avg = {k: v.clone() for k, v in ema.module.state_dict().items()}
    for k in avg:
        assert torch.equal(avg[k], O.ema_update(p0[k], p1[k], 0.9, False)), k
    save("ema.npz", **{"p0." + k: v for k, v in p0.items()}, **{"p1." + k: v for k, v in p1.items()},
         **{"avg." + k: v for k, v in avg.items()})

    # ---- (7) train step: loss + grads for a fixed (ts, noise, mask) triple -------------------
    kind, cfg = CONFIGS["nu_like"]
    with torch.enable_grad():
        model = UNet1D(**cfg)
        sdm = make_state_dict({k: v.shape for k, v in model.state_dict().items()}, seed=1234)
        model.load_state_dict(sdm)
        g = torch.Generator().manual_seed(21)
        Bt = 64
        y = torch.rand(Bt, cfg["input_dim"], generator=g)
        cond = torch.rand(Bt, cfg["cond_dim"], generator=g)
        ts = torch.randint(0, T, (1, Bt), generator=g)
        noise = torch.randn(Bt, cfg["input_dim"], generator=g)
        mask = (torch.rand(Bt, 1, generator=g) < 0.9).float()
        y_t = torch.squeeze(bufs["sqrt_alphas_cumprod"][ts, None] * y + bufs["sqrt_one_minus_alphas_cumprod"][ts, None] * noise)
        loss = torch.nn.functional.mse_loss(noise, model(y_t, ts / T, cond, mask))
        loss.backward()
        full = {"model." + k: v for k, v in sdm.items()}
        full.update(bufs)
        assert torch.equal(loss.detach(), O.q_sample_loss(full, T, y, cond, ts, noise, mask))
        grads = {"grad." + k: p.grad.clone() for k, p in model.named_parameters()}
    save("train_step.npz", y=y, cond=cond, ts=ts, noise=noise, mask=mask, loss=loss.detach(), **grads)

    (GOLD / "state_dict_manifest.json").write_text(json.dumps(manifest, indent=0, sort_keys=True))
    print("all oracle-vs-reference assertions passed")


if __name__ == "__main__":
    main()
