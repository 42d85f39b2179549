"""Import shim: expose this package under the reference's module names.

After `install_reference_aliases()`,
    from ddpm_opt.UNetCF import UNet1D
    from ddpm_opt.classifier_free_NU import DDPM, nu_data_load, custom_decoder, rate_calc
    from ddpm_opt.diffusion import generate_cosine_schedule, init_weights
    from ddpm_opt.ema import ExponentialMovingAverage
resolve to the kernel-backed implementations, so the reference's scripts, trajectory
generators (datasets/*_trajectory_gen.py) and baselines import them unchanged.
"""
from __future__ import annotations

import sys
import types


def install_reference_aliases(force: bool = False):
    from . import co, ema, msr, nu, schedule, unet
    if "ddpm_opt" in sys.modules and not force and not getattr(sys.modules["ddpm_opt"], "_diffsg_b200", False):
        raise RuntimeError("a different 'ddpm_opt' package is already imported; pass force=True to shadow it")
    pkg = types.ModuleType("ddpm_opt")
    pkg.__path__ = []
    pkg._diffsg_b200 = True
    table = {"UNetCF": unet, "ema": ema, "diffusion": schedule, "classifier_free_MSR": msr,
             "classifier_free_NU": nu, "classifier_free_CO": co}
    sys.modules["ddpm_opt"] = pkg
    for name, mod in table.items():
        sys.modules[f"ddpm_opt.{name}"] = mod
        setattr(pkg, name, mod)
    return pkg
