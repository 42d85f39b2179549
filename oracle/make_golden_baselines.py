#!/usr/bin/env python
"""Golden vectors for the batched baseline forwards: the reference's own MTFNN / PPOAgent classes
(baselines/MTFNN.py, baselines/PPO.py, imported UNMODIFIED from /root/reference) carrying the bundled checkpoints
(ckpts/mtfnn_*.pt, ckpts/ppo_*.pt), evaluated with torch on the CPU on seeded inputs.  Stored with the checkpoint tensors
in tests/golden/baselines.npz so the GPU test needs nothing from /root/reference.

    python oracle/make_golden_baselines.py      # test infrastructure only
"""
import sys
from collections import OrderedDict
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parents[1] / "tests" / "golden" / "baselines.npz"

if __name__ == "__main__":
    sys.path.insert(0, str(REF))
    from baselines.MTFNN import MTFNN as RefMTFNN
    from baselines.PPO import PPOAgent as RefPPO
    out = {}
    g = torch.Generator().manual_seed(2025)

    def seq(dims, head):     # the Sequential models the reference builds inline (MTFNN.py:43-52, 122-131)
        d = OrderedDict()
        for i in range(len(dims) - 1):
            d[f"lin{i + 1}"] = nn.Linear(dims[i], dims[i + 1])
            d[f"act{i + 1}"] = nn.ReLU() if i < len(dims) - 2 else head
        return nn.Sequential(d)

    cases = {"mtfnn_co": lambda sd: seq([sd["lin1.weight"].shape[1], 32, 64, 16, sd["lin4.weight"].shape[0]], nn.Sigmoid()),
             "mtfnn_msr_3c": lambda sd: seq([sd["lin1.weight"].shape[1], 8, 16, 8, sd["lin4.weight"].shape[0]], nn.Softmax(dim=1)),
             "mtfnn_msr_80c": lambda sd: seq([sd["lin1.weight"].shape[1], 8, 16, 8, sd["lin4.weight"].shape[0]], nn.Softmax(dim=1)),
             "mtfnn_nu": lambda sd: RefMTFNN(sd["lin1.weight"].shape[1], sd["lin5.weight"].shape[0]),
             "ppo_co": lambda sd: RefPPO(sd["actor.0.weight"].shape[1], sd["actor.6.weight"].shape[0]),
             "ppo_msr_3c": lambda sd: RefPPO(sd["actor.0.weight"].shape[1], sd["actor.6.weight"].shape[0]),
             "ppo_msr_80c": lambda sd: RefPPO(sd["actor.0.weight"].shape[1], sd["actor.6.weight"].shape[0]),
             "ppo_nu": lambda sd: RefPPO(sd["actor.0.weight"].shape[1], sd["actor.6.weight"].shape[0])}
    for name, build in cases.items():
        sd = torch.load(REF / "ckpts" / f"{name}.pt", map_location="cpu")
        model = build(sd)
        model.load_state_dict(sd)
        model.eval()
        in_dim = (sd["lin1.weight"] if "lin1.weight" in sd else sd["actor.0.weight"]).shape[1]
        x = torch.rand(300, in_dim, generator=g)
        with torch.no_grad():
            if name.startswith("ppo"):
                value, dist = model(x)
                out[f"{name}.value"], out[f"{name}.mu"], out[f"{name}.std"] = value.numpy(), dist.mean.numpy(), dist.stddev.numpy()
            else:
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")          # the reference's nn.Softmax() has no dim (legacy: dim=1 for 2-D)
                    out[f"{name}.y"] = model(x.clone()).numpy()
        out[f"{name}.x"] = x.numpy()
        for k, v in sd.items():
            out[f"{name}.sd.{k}"] = v.numpy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, OUT.stat().st_size, "bytes;", len(out), "arrays")
