#!/usr/bin/env python
"""Throughput of sample() for one network shape under the tensor-core build variant given by
DIFFSG_TC_VARIANT (the library is rebuilt for it first).  Also checks the result against the fp32 engine.
Usage: DIFFSG_TC_VARIANT=chunk=32,aslots=3,region=64,ctas=3 tools/variant_probe.py co 262144"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import diffsg_b200 as D
from diffsg_b200 import _lib

NETS = {
    "co": dict(input_dim=3, proj_dim=64, cond_dim=9, dims=(64, 32, 16, 8), is_attn=(False,) * 4, middle_attn=False, n_blocks=3),
    "nu": dict(input_dim=5, proj_dim=32, cond_dim=6, dims=(32, 16, 8), is_attn=(False,) * 3, middle_attn=False, n_blocks=2),
    "msr80c": dict(input_dim=80, proj_dim=128, cond_dim=80, dims=(64, 32, 16, 8), is_attn=(False,) * 4, middle_attn=False, n_blocks=2),
}


def build(net, precision):
    torch.manual_seed(0)
    model = D.UNet1D(**net)
    model.precision = precision
    T = 20
    alphas = 1.0 - D.generate_cosine_schedule(T)
    ddpm = D.msr.DDPM(T, model, net["input_dim"], 20.0, alphas, "cuda", (1, net["input_dim"]),
                      {"scaler_min": 0.5, "scaler_max": 2.5, "W": 20.0}, 0.1, 0.9999, 10, 5, False)
    ddpm.apply(D.init_weights)
    return ddpm.to("cuda")


def main():
    name, rows = sys.argv[1], int(sys.argv[2])
    _lib.build_library(force=True)
    net = NETS[name]
    ddpm = build(net, "fp16x2")
    ddpm.noise_mode = "philox"
    cond = torch.rand(rows, net["cond_dim"], device="cuda")
    for _ in range(3):
        ddpm.philox_offset = 0
        y = ddpm.sample(cond, omega=500.0)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(3):
        ddpm.philox_offset = 0
        y = ddpm.sample(cond, omega=500.0)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 3
    # parity of the variant: one forward against the exact-fp32 engine on the same weights
    ref = build(net, "fp32")
    ref.load_state_dict(ddpm.state_dict())
    n = 4096
    x = torch.randn(n, net["input_dim"], device="cuda")
    ts = torch.randint(0, 20, (1, n), device="cuda")
    mask = torch.ones(n, 1, device="cuda")
    with torch.no_grad():
        a = ddpm.model(x, ts / 20, cond[:n], mask)
        b = ref.model(x, ts / 20, cond[:n], mask)
    err = float((a - b).norm() / b.norm())
    print(f"{name} variant={_lib.TC_VARIANT} info={ddpm.model.engine().info()} rows={rows}: {rows / ms * 1e3:,.0f} solutions/s "
          f"({ms:.2f} ms), forward rel-L2 vs fp32 engine {err:.2e}, finite={bool(torch.isfinite(y).all())}", flush=True)


if __name__ == "__main__":
    main()
