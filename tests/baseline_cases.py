"""Shared by the CPU and GPU baseline tests: the mirror modules of diffsg_b200.baselines carrying the reference
checkpoints stored in tests/golden/baselines.npz."""
from pathlib import Path

import numpy as np
import torch

from diffsg_b200 import baselines as B

CASES = {"mtfnn_co": "seq_co", "mtfnn_msr_3c": "seq_msr", "mtfnn_msr_80c": "seq_msr", "mtfnn_nu": "nu",
         "ppo_co": "ppo", "ppo_msr_3c": "ppo", "ppo_msr_80c": "ppo", "ppo_nu": "ppo"}


def golden():
    return np.load(Path(__file__).parent / "golden" / "baselines.npz")


def state_dict(name, z):
    pre = f"{name}.sd."
    return {k[len(pre):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(pre)}


def build_case(name, z):
    sd = state_dict(name, z)
    kind = CASES[name]
    if kind == "ppo":
        m = B.PPOAgent(sd["actor.0.weight"].shape[1], sd["actor.6.weight"].shape[0])
    elif kind == "nu":
        m = B.MTFNN(sd["lin1.weight"].shape[1], sd["lin5.weight"].shape[0])
    else:
        mk = B.mtfnn_co_model if kind == "seq_co" else B.mtfnn_msr_model
        m = mk(sd["lin1.weight"].shape[1], sd["lin4.weight"].shape[0])
    m.load_state_dict(sd)          # strict: the reference's checkpoint keys and shapes
    return m
