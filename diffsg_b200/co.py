"""Computation-offloading (CO) front-end: drop-in for ddpm_opt/classifier_free_CO.py."""
from __future__ import annotations

import numpy as np
import torch

from . import objectives
from .ddpm import DDPMBase
from .ema import ExponentialMovingAverage  # noqa: F401
from .schedule import generate_cosine_schedule, init_weights  # noqa: F401
from .unet import UNet1D  # noqa: F401

# constants the reference appends to every row (CO.py:174-180): F_t, kappa, Pt, PI, theta, B, N0
CO_COMMON = (2.5e9, 1e-28, 0.3, 0.1, 1.0, 10e5, 7.96159e-13)


class DDPM(DDPMBase):
    """Constructor signature of reference classifier_free_CO.py:60-72."""

    def __init__(self, T, model, node_num, alphas, device, data_size, custom_config=None, uncond_prob=0.1,
                 ema_decay=0.9999, ema_start=1000, ema_update_rate=5, debug=False):
        super().__init__()
        self.node_num = node_num
        self._setup(T, model, alphas, device, data_size, custom_config, uncond_prob, ema_decay, ema_start,
                    ema_update_rate, debug)

    def _decode_record(self, j, y):
        return customized_real_decoder(y)


def customized_real_decoder(Y_pred):
    """Row softmax, zeroed where every logit < -10 (reference CO.py:281-290)."""
    return objectives.co_decode(Y_pred).reshape(Y_pred.shape)


def cost_calc(X, Y):
    """Total cost of an allocation (reference CO.py:255-278; generalised from 3 nodes to n)."""
    return objectives.co_cost(X, Y)


def data_preprocess_co(X):
    """6 raw features/node + 7 constants -> 3 costs/node [local, offload transition, ideal
    offload execution] (reference utils/dataset.py:26-51)."""
    n = (X.shape[1] - 7) // 6
    F_t, kappa, Pt, PI, _theta, Bw, N0 = (X[:, -7 + i] for i in range(7))
    f = X[:, :6 * n].reshape(X.shape[0], n, 6)
    data, cyc, freq, h, off = f[..., 0], f[..., 1], f[..., 2], f[..., 3], f[..., 4]
    interf = (Pt[:, None] * h ** 2).sum(axis=1)
    sinr = Pt[:, None] * h ** 2 / (N0 + interf)[:, None]
    r_u = Bw[:, None] * np.log2(1.0 + sinr)
    out = np.zeros((X.shape[0], n, 3))
    out[..., 0] = off * cyc / freq + (1.0 - off) * kappa[:, None] * freq ** 2 * cyc
    out[..., 1] = off * data / r_u + (1.0 - off) * Pt[:, None] * data / r_u
    out[..., 2] = off * cyc / F_t[:, None] + (1.0 - off) * PI[:, None] * cyc / F_t[:, None]
    return out.reshape(X.shape[0], 3 * n)


def co_data_load(dataset_path):
    """CSV rows `6 features x node | class | alloc[node]` (reference CO.py:158-200)."""
    import pandas as pd
    src = np.array(pd.read_csv(dataset_path, header=None), dtype=np.float64)
    n = (src.shape[1] - 1) // 7
    X, Y = src[:, :6 * n], src[:, -n:]
    X = np.concatenate((X, np.tile(np.array(CO_COMMON)[None, :], (X.shape[0], 1))), axis=1)
    X = data_preprocess_co(X)
    keep = np.all(X < 10.0, axis=1)
    X, Y = X[keep], Y[keep]
    lo, hi = np.min(X), np.max(X)
    X = (X - lo) / (hi - lo)
    cfg = {"sfn": 3, "cfn": 0, "cdim": 1, "scaler_min": lo, "scaler_max": hi}
    n_tr, n_te = int(src.shape[0] * 0.7), int(src.shape[0] * 0.3)
    return X[:n_tr], Y[:n_tr], X[-n_te:], Y[-n_te:], cfg


@torch.no_grad()
def evaluate(diffusion_model, X_test, Y_test, custom_config, omega=500.0, batch_size=512):
    """`load_test_co` core (reference CO.py:315-356): exceeded ratio, accuracy, terrible count."""
    dev = diffusion_model.betas.device
    X = torch.as_tensor(X_test, dtype=torch.float32, device=dev)
    Y = torch.as_tensor(Y_test, dtype=torch.float32, device=dev)
    Y_pred = torch.cat([diffusion_model.sample(X[i:i + batch_size], omega).reshape(-1, Y.shape[1])
                        for i in range(0, X.shape[0], batch_size)])
    lo, hi = custom_config["scaler_min"], custom_config["scaler_max"]
    Xs = X * (hi - lo) + lo
    dec = customized_real_decoder(Y_pred)
    pred_cost, true_cost = cost_calc(Xs, dec), cost_calc(Xs, Y)
    same = ((dec > 0.1) == (Y > 0.1)).all(dim=1)
    terrible = (pred_cost / true_cost > 1.2) & (pred_cost > 10.0)
    return dict(exceeded_ratio=float(pred_cost.sum() / true_cost.sum()),
                avg_cost_diff=float((pred_cost - true_cost).mean()), accuracy=int(same.sum()),
                terrible=int(terrible.sum()), pred_cost=pred_cost, true_cost=true_cost, Y_pred=Y_pred)


# ---- script-level entry points (reference classifier_free_CO.py:203-252, 293-356): same names, same constants
CO_NET = dict(proj_dim=64, dims=(64, 32, 16, 8), is_attn=(False, False, False, False), middle_attn=False, n_blocks=3)


def _co_net(n, sfn=3):
    return dict(input_dim=n, cond_dim=sfn * n, **CO_NET)


def train_ddpm_co(dataset_path="../datasets/3nodes_50000samples_new.csv", epochs=200, lr=0.005, milestones=(15, 80, 150),
                  use_ema=False, device=None, **fit_kw):
    """`train_ddpm_co()` of the reference (T = 20, Adam lr 0.005, MultiStepLR [15, 80, 150], bs 512, 200 epochs)."""
    from . import scripts
    X_train, Y_train, _, _, cfg = co_data_load(dataset_path)
    n = Y_train.shape[1]
    return scripts.train(DDPM, (n,), _co_net(n, cfg["sfn"]), cfg, X_train, Y_train, epochs=epochs, lr=lr,
                         milestones=milestones, use_ema=use_ema, device=device, **fit_kw)


@torch.no_grad()
def load_test_co(ckpt_path, dataset_path="../datasets/3nodes_50000samples_new.csv", omega=500.0, device=None, verbose=True):
    """`load_test_co(ckpt_path)` of the reference; also returns the numbers it prints."""
    from . import scripts
    X_train, Y_train, X_test, Y_test, cfg = co_data_load(dataset_path)
    n = Y_train.shape[1]
    ddpm = scripts.load(DDPM, (n,), _co_net(n, cfg["sfn"]), cfg, ckpt_path, device=device)
    out = evaluate(ddpm, X_test, Y_test, cfg, omega=omega, batch_size=512)
    if verbose:
        scripts.report([("Y_pred", customized_real_decoder(out["Y_pred"])), ("Y_test", Y_test),
                        ("pred_cost", out["pred_cost"]), ("true_cost", out["true_cost"])],
                       [f"exceeded ratio: {out['exceeded_ratio']}", f"avg cost diff:\n {out['avg_cost_diff']}",
                        f"terrible samples num: {out['terrible']}/{len(X_test)}.", f"accuracy: {out['accuracy']}/{len(X_test)}"])
    return out
