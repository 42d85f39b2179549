import json
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
GOLD = ROOT / "tests" / "golden"
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_golden(name):
    with np.load(GOLD / name) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def manifest():
    return json.loads((GOLD / "state_dict_manifest.json").read_text())


def rel_l2(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).reshape(-1)
    b = torch.as_tensor(b, dtype=torch.float64).reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def standin_model(name, device="cpu"):
    """diffsg_b200 UNet1D + DDPM carrying the deterministic stand-in weights of oracle/standin.py."""
    import diffsg_b200 as D
    from oracle.standin import CONFIGS, make_state_dict
    kind, cfg = CONFIGS[name]
    model = D.UNet1D(**cfg)
    sd = make_state_dict({k: v.shape for k, v in model.state_dict().items()}, seed=1234)
    model.load_state_dict(sd)
    alphas = 1.0 - D.generate_cosine_schedule(20)
    M = cfg["input_dim"]
    if kind == "nu":
        ddpm = D.nu.DDPM(20, model, 3, 18.0, alphas, device, (1, M), {"width": 400, "height": 400}, 0.1, 0.9999, 10, 5, False)
    elif kind == "co":
        ddpm = D.co.DDPM(20, model, M, alphas, device, (1, M), {}, 0.1, 0.9999, 10, 5, False)
    else:
        ddpm = D.msr.DDPM(20, model, M, 10.0, alphas, device, (1, M), {}, 0.1, 0.9999, 10, 5, False)
    return ddpm.to(device), cfg


def nu_checkpoint_model(device="cpu"):
    import diffsg_b200 as D
    ck = load_golden("nu_ckpt.npz")
    model = D.UNet1D(input_dim=5, proj_dim=32, cond_dim=6, dims=(32, 16, 8), is_attn=(False,) * 3,
                     middle_attn=False, n_blocks=2)
    alphas = 1.0 - D.generate_cosine_schedule(20)
    ddpm = D.nu.DDPM(20, model, 3, 18.0, alphas, device, (1, 5), {"width": 400, "height": 400}, 0.1, 0.9999, 10, 5, False)
    sd = {k: torch.tensor(v) for k, v in ck.items()}
    missing = ddpm.load_state_dict(sd, strict=False)
    assert all(k.startswith("ema.") for k in missing.missing_keys) and not missing.unexpected_keys
    return ddpm.to(device)
