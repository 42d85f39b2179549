"""NOMA-UAV (NU) front-end: drop-in for ddpm_opt/classifier_free_NU.py."""
from __future__ import annotations

import numpy as np
import torch

from . import objectives
from .ddpm import DDPMBase
from .ema import ExponentialMovingAverage  # noqa: F401
from .msr import parse_scalar_from_name
from .schedule import generate_cosine_schedule, init_weights  # noqa: F401
from .unet import UNet1D  # noqa: F401


class DDPM(DDPMBase):
    """Constructor signature of reference classifier_free_NU.py:84-97."""

    def __init__(self, T, model, K, P_sum, alphas, device, data_size, custom_config=None, uncond_prob=0.1,
                 ema_decay=0.9999, ema_start=1000, ema_update_rate=5, debug=False):
        super().__init__()
        self.K, self.P_sum = K, P_sum
        self._setup(T, model, alphas, device, data_size, custom_config, uncond_prob, ema_decay, ema_start,
                    ema_update_rate, debug)

    def _decode_record(self, j, y):
        cfg = self.custom_config or {}
        return custom_decoder(y, cfg.get("width", 400), cfg.get("height", 400), self.P_sum)


def custom_decoder(Y_pred, width, height, P_sum):
    """UAV position: global min-max of the first two columns scaled to the area; powers:
    P_sum * softmax (reference NU.py:267-276)."""
    return objectives.nu_decode(Y_pred, width, height, P_sum)


def rate_calc(Y_pred_decoded, X):
    """NOMA sum rate with SIC ordered by channel gain (reference NU.py:279-303), vectorised
    on the GPU instead of the reference's per-row Python loop."""
    return objectives.nu_rate(Y_pred_decoded, X)


def nu_data_load(dataset_path, width, height):
    """CSV rows `x1,y1,..,xK,yK | u_x,u_y | p1..pK | rate` (reference NU.py:184-210)."""
    import pandas as pd
    src = np.array(pd.read_csv(dataset_path, header=None), dtype=np.float64)
    K = (src.shape[1] - 3) // 3
    P_sum = parse_scalar_from_name(dataset_path, "mw")
    X, Y, R = src[:, :2 * K].copy(), src[:, 2 * K:2 + 3 * K].copy(), src[:, -1]
    X[:, 0::2] /= width
    X[:, 1::2] /= height
    Y[:, 0] /= width
    Y[:, 1] /= height
    Y[:, 2:] /= P_sum
    cfg = {"K": K, "P_sum": P_sum, "cdim": 1, "width": width, "height": height}
    n_tr, n_te = int(src.shape[0] * 0.7), int(src.shape[0] * 0.3)
    return X[:n_tr], Y[:n_tr], X[-n_te:], Y[-n_te:], R[-n_te:], cfg


@torch.no_grad()
def evaluate(diffusion_model, X_test, Y_test, custom_config, omega=500, batch_size=512):
    """`load_test_nu` core (reference NU.py:331-361). Returns dict(less_ratio, ...)."""
    dev = diffusion_model.betas.device
    width, height, P_sum = custom_config["width"], custom_config["height"], custom_config["P_sum"]
    X = torch.as_tensor(X_test, dtype=torch.float32, device=dev)
    Y = torch.as_tensor(Y_test, dtype=torch.float32, device=dev).clone()
    Y_pred = torch.cat([diffusion_model.sample(X[i:i + batch_size], omega).reshape(-1, Y.shape[1])
                        for i in range(0, X.shape[0], batch_size)])
    Xs = X.clone()
    Xs[:, 0::2] *= width
    Xs[:, 1::2] *= height
    Y[:, 0] *= width
    Y[:, 1] *= height
    Y[:, 2:] *= P_sum
    pred_rate = rate_calc(custom_decoder(Y_pred, width, height, P_sum), Xs)
    true_rate = rate_calc(Y, Xs)
    return dict(less_ratio=float(pred_rate.sum() / true_rate.sum()),
                avg_rate_diff=float((pred_rate - true_rate).mean()), pred_rate=pred_rate,
                true_rate=true_rate, Y_pred=Y_pred)


# ---- script-level entry points (reference classifier_free_NU.py:213-264, 306-361): same names, same constants
NU_NET = dict(proj_dim=32, dims=(32, 16, 8), is_attn=(False, False, False), middle_attn=False, n_blocks=2)


def _nu_net(K):
    return dict(input_dim=2 + K, cond_dim=2 * K, **NU_NET)


def train_ddpm_nu(dataset_path="../datasets/3u_18mW_10000samples.csv", width=400, height=400, epochs=200, lr=0.004,
                  milestones=(80, 200), use_ema=False, device=None, **fit_kw):
    """`train_ddpm_nu()` of the reference (T = 20, Adam lr 0.004, MultiStepLR [80, 200], bs 512, 200 epochs)."""
    from . import scripts
    X_train, Y_train, _, _, _, cfg = nu_data_load(dataset_path, width, height)
    K, P_sum = cfg["K"], cfg["P_sum"]
    return scripts.train(DDPM, (K, P_sum), _nu_net(K), cfg, X_train, Y_train, epochs=epochs, lr=lr,
                         milestones=milestones, use_ema=use_ema, device=device, **fit_kw)


@torch.no_grad()
def load_test_nu(ckpt_path, dataset_path="../datasets/3u_18mW_10000samples.csv", width=400, height=400, omega=500,
                 device=None, verbose=True):
    """`load_test_nu(ckpt_path)` of the reference; also returns the numbers it prints."""
    from . import scripts
    _, _, X_test, Y_test, _, cfg = nu_data_load(dataset_path, width, height)
    K, P_sum = cfg["K"], cfg["P_sum"]
    ddpm = scripts.load(DDPM, (K, P_sum), _nu_net(K), cfg, ckpt_path, device=device)
    out = evaluate(ddpm, X_test, Y_test, cfg, omega=omega, batch_size=512)
    if verbose:
        scripts.report([("Y_pred", custom_decoder(out["Y_pred"], width, height, P_sum)), ("pred_rate", out["pred_rate"]),
                        ("true_rate", out["true_rate"])],
                       [f"less ratio: {out['less_ratio']}", f"avg rate diff:\n {out['avg_rate_diff']}"], precision=8)
    return out
