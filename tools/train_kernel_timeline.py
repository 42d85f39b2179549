#!/usr/bin/env python
"""Per-launch timeline of one eager training step (80c net): kernel, grid, dynamic smem, duration, in launch order."""
import json
import sys
import tempfile
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
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
    with tempfile.TemporaryDirectory() as d:
        path = Path(d) / "trace.json"
        prof.export_chrome_trace(str(path))
        tr_ = json.loads(path.read_text())
    evs = [e for e in tr_["traceEvents"] if e.get("cat") == "kernel"]
    evs.sort(key=lambda e: e["ts"])
    for e in evs:
        a = e.get("args", {})
        name = e["name"].split("(")[0].split("::")[-1][:28]
        print(f"{name:28s} grid={str(a.get('grid')):16s} smem={a.get('shared memory', 0):6} regs={a.get('registers per thread', 0):4} {e['dur']:8.1f} us")
