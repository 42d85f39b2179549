"""torchrun smoke test of the data-parallel trainer over NCCL (one rank per GPU).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_train_smoke.py
"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffsg_b200 as D  # noqa: E402
from diffsg_b200.parallel import DataParallelTrainer, shard_rows  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)                                   # identical init on every rank
net = dict(input_dim=80, proj_dim=128, cond_dim=80, dims=(64, 32, 16, 8), is_attn=(False,) * 4, middle_attn=False, n_blocks=2)
model = D.UNet1D(**net)
ddpm = D.msr.DDPM(20, model, 80, 20.0, 1.0 - D.generate_cosine_schedule(20), dev, (1, 80), {}).to(dev)
ddpm.apply(D.init_weights)
GRAPH = os.environ.get("CUDA_GRAPH", "0") == "1"
tr = DataParallelTrainer(ddpm, lr=1e-3, cuda_graph=GRAPH)
B = int(os.environ.get("PER_GPU_BATCH", "8192"))
g = torch.Generator().manual_seed(1)
X = torch.rand(B * world, 80, generator=g)
Y = torch.rand(B * world, 80, generator=g) * 0.5
sl = shard_rows(B * world, rank, world)
x, y = X[sl].to(dev), Y[sl].to(dev)
torch.manual_seed(100 + rank)                          # per-rank RNG stream for (ts, noise, mask)
losses = []
STEPS = int(os.environ.get("STEPS", "50"))
for i in range(10 + STEPS):                            # 10 warm-up steps, then STEPS timed ones
    if i == 10:
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    losses.append(tr.step(y, x))                       # device scalar: no host sync inside the timed region
torch.cuda.synchronize(); dist.barrier(); dt = (time.perf_counter() - t0) / STEPS
losses = [float(v) for v in losses]
chk = tr.flat.flat.double().sum().reshape(1)
allc = [torch.empty_like(chk) for _ in range(world)]
dist.all_gather(allc, chk)
same = all(torch.equal(allc[0], c) for c in allc)
if rank == 0:
    print(f"dp_train_smoke: world={world} cuda_graph={GRAPH} per_gpu_batch={B} loss {losses[0]:.4f} -> {losses[-1]:.4f}, "
          f"{dt * 1e3:.1f} ms/step, {B * world / dt:.0f} samples/s, replicas identical: {same}")
assert same and losses[-1] < losses[0]
dist.destroy_process_group()
