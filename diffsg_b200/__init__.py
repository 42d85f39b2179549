"""diffsg_b200 — B200-native (sm_100a) CFG-DDPM solver path of qiyu3816/DiffSG.

Public surface (mirrors the reference's Python API for this path):
    UNet1D, ExponentialMovingAverage, generate_cosine_schedule, init_weights,
    msr.DDPM / nu.DDPM / co.DDPM (+ loaders, decoders, objectives of each script).
`install_reference_aliases()` makes `from ddpm_opt.classifier_free_MSR import DDPM, ...`
resolve to these modules so the reference scripts run unchanged.
"""
from .ema import ExponentialMovingAverage
from .schedule import generate_cosine_schedule, generate_linear_schedule, init_weights
from .unet import UNet1D, infer_config_from_state_dict
from . import msr, nu, co, objectives  # noqa: E402
from .compat import install_reference_aliases

__all__ = ["UNet1D", "ExponentialMovingAverage", "generate_cosine_schedule", "generate_linear_schedule",
           "init_weights", "infer_config_from_state_dict", "msr", "nu", "co", "objectives",
           "install_reference_aliases"]
__version__ = "0.1.0"
