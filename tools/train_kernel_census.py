#!/usr/bin/env python
"""Kernel census of one eager eps-MSE training step of the 80c net (which launches make up the ~1 300 of a step)."""
import collections
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
import diffsg_b200 as D  # noqa: E402
from diffsg_b200.parallel import DataParallelTrainer  # noqa: E402

if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    dev = torch.device("cuda:0")
    ddpm = bench.build_model(dev)
    tr = DataParallelTrainer(ddpm, lr=1e-3, cuda_graph=False)
    y = torch.rand(B, bench.NET["input_dim"], device=dev)
    c = torch.rand(B, bench.NET["cond_dim"], device=dev)
    for _ in range(3):
        tr.step(y, c)
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        tr.step(y, c)
        torch.cuda.synchronize()
    cnt, tim = collections.Counter(), collections.Counter()
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            name = ev.name[:70]
            cnt[name] += 1
            tim[name] += ev.device_time
    print("kernels per step:", sum(cnt.values()), " device us:", sum(tim.values()))
    for name, n in cnt.most_common(25):
        print(f"{n:5d} {tim[name]:9.1f} us  {name}")
