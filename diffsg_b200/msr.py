"""Maximum-sum-rate (MSR) front-end: drop-in for ddpm_opt/classifier_free_MSR.py."""
from __future__ import annotations

import re

import numpy as np
import torch

from . import objectives
from .ddpm import DDPMBase
from .ema import ExponentialMovingAverage  # noqa: F401  (re-exported like the reference script)
from .schedule import generate_cosine_schedule, init_weights  # noqa: F401
from .unet import UNet1D  # noqa: F401


class DDPM(DDPMBase):
    """Constructor signature of reference classifier_free_MSR.py:55-68."""

    def __init__(self, T, model, M, W, alphas, device, data_size, custom_config=None, uncond_prob=0.1,
                 ema_decay=0.9999, ema_start=1000, ema_update_rate=5, debug=False):
        super().__init__()
        self.M, self.W = M, W
        self._setup(T, model, alphas, device, data_size, custom_config, uncond_prob, ema_decay, ema_start,
                    ema_update_rate, debug)

    def _decode_record(self, j, y):
        # reference MSR.py:143-149: plain softmax for the first three records, decoder afterwards
        if j <= 2:
            return torch.softmax(y, dim=1)
        return custom_decoder(y)


def custom_decoder(Y_pred):
    """softmax_row((Y - min) / (max - min)) with GLOBAL min/max (reference MSR.py:239-245)."""
    ones = torch.ones_like(Y_pred)
    _, p = objectives.msr_decode_rate(Y_pred, ones, 1.0, return_alloc=True)
    return p.reshape(Y_pred.shape)


def parse_scalar_from_name(path: str, unit: str) -> float:
    """`3c_10w_10000samples.csv` -> 10.0 (unit 'w'); also accepts `*_ood.csv` names, which the
    reference's `split('_')[-2][:-1]` parser cannot."""
    name = str(path).replace("\\", "/").split("/")[-1]
    m = re.search(r"_(\d+(?:\.\d+)?)" + re.escape(unit) + r"_", name, flags=re.IGNORECASE)
    if not m:
        raise ValueError(f"cannot parse '<number>{unit}' from dataset name {name!r}")
    return float(m.group(1))


def msr_data_load(dataset_path):
    """CSV rows `g[M] | rate | p*[M]` -> min-max scaled conditions, 70/30 head/tail split
    (reference MSR.py:159-184)."""
    import pandas as pd
    src = np.array(pd.read_csv(dataset_path, header=None))
    M = (src.shape[1] - 1) // 2
    W = parse_scalar_from_name(dataset_path, "w")
    X, Y = src[:, :M], src[:, -M:]
    lo, hi = np.min(X), np.max(X)
    X = (X - lo) / (hi - lo)
    cfg = {"M": M, "W": W, "sfn": 1, "cfn": 0, "cdim": 1, "scaler_min": lo, "scaler_max": hi}
    n_tr, n_te = int(src.shape[0] * 0.7), int(src.shape[0] * 0.3)
    return X[:n_tr], Y[:n_tr], X[-n_te:], Y[-n_te:], cfg


@torch.no_grad()
def evaluate(diffusion_model, X_test, Y_test, custom_config, omega=500, batch_size=512):
    """`load_test_msr` core (reference MSR.py:272-298) with the decode + rate on the GPU.
    Returns dict(less_ratio, avg_rate_diff, pred_rate, true_rate, Y_pred)."""
    dev = diffusion_model.betas.device
    X = torch.as_tensor(X_test, dtype=torch.float32, device=dev)
    Y = torch.as_tensor(Y_test, dtype=torch.float32, device=dev)
    Y_pred = torch.cat([diffusion_model.sample(X[i:i + batch_size], omega).reshape(-1, X.shape[1])
                        for i in range(0, X.shape[0], batch_size)])
    lo, hi = custom_config["scaler_min"], custom_config["scaler_max"]
    g = X * (hi - lo) + lo
    pred_rate = objectives.msr_decode_rate(Y_pred, g, custom_config["W"])
    true_rate = objectives.msr_rate(Y, g)
    return dict(less_ratio=float(pred_rate.sum() / true_rate.sum()),
                avg_rate_diff=float((pred_rate - true_rate).mean()), pred_rate=pred_rate,
                true_rate=true_rate, Y_pred=Y_pred)


# ---- script-level entry points (reference classifier_free_MSR.py:187-236, 248-298): same names, same constants
MSR_NET = dict(proj_dim=128, dims=(64, 32, 16, 8), is_attn=(False, False, False, False), middle_attn=False, n_blocks=2)


def _msr_net(M, sfn=1):
    return dict(input_dim=M, cond_dim=sfn * M, **MSR_NET)


def train_ddpm_msr(dataset_path="../datasets/3c_10w_10000samples.csv", epochs=200, lr=0.005, milestones=(100, 150),
                   use_ema=False, device=None, **fit_kw):
    """`train_ddpm_msr()` of the reference (T = 20, Adam lr 0.005, MultiStepLR [100, 150], bs 512, 200 epochs)."""
    from . import scripts
    X_train, Y_train, _, _, cfg = msr_data_load(dataset_path)
    M, W = cfg["M"], cfg["W"]
    return scripts.train(DDPM, (M, W), _msr_net(M, cfg["sfn"]), cfg, X_train, Y_train, epochs=epochs, lr=lr,
                         milestones=milestones, use_ema=use_ema, device=device, **fit_kw)


@torch.no_grad()
def load_test_msr(ckpt_path, dataset_path="../datasets/3c_10w_10000samples.csv", omega=500, device=None, verbose=True):
    """`load_test_msr(ckpt_path)` of the reference: sample the test split in 512-row batches at omega = 500, decode,
    print the same report; additionally returns the numbers (the reference returns None)."""
    from . import scripts
    _, _, X_test, Y_test, cfg = msr_data_load(dataset_path)
    M, W = cfg["M"], cfg["W"]
    ddpm = scripts.load(DDPM, (M, W), _msr_net(M, cfg["sfn"]), cfg, ckpt_path, device=device)
    out = evaluate(ddpm, X_test, Y_test, cfg, omega=omega, batch_size=512)
    if verbose:
        lo, hi = cfg["scaler_min"], cfg["scaler_max"]
        Xs = torch.as_tensor(X_test, dtype=torch.float32, device=out["Y_pred"].device) * (hi - lo) + lo
        scripts.report([("Y_pred", W * custom_decoder(out["Y_pred"])), ("Y_test", Y_test), ("X_test", Xs),
                        ("pred_rate", out["pred_rate"]), ("true_rate", out["true_rate"])],
                       [f"less ratio: {out['less_ratio']}", f"avg rate diff:\n {out['avg_rate_diff']}"])
    return out
