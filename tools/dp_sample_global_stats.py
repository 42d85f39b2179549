"""torchrun check of sharded sampling with whole-batch statistics over NCCL: every rank samples its row shard of
ONE batch with `stats_group`, rank 0 compares the gathered result with the un-sharded call.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_sample_global_stats.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffsg_b200 as D  # noqa: E402
from diffsg_b200.parallel import shard_rows  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
net = dict(input_dim=80, proj_dim=128, cond_dim=80, dims=(64, 32, 16, 8), is_attn=(False,) * 4, middle_attn=False, n_blocks=2)
T = 20
ddpm = D.msr.DDPM(T, D.UNet1D(**net), 80, 20.0, 1.0 - D.generate_cosine_schedule(T), dev, (1, 80), {}).to(dev)
ddpm.apply(D.init_weights)
B = int(os.environ.get("ROWS", "20001"))                # ragged on purpose
g = torch.Generator().manual_seed(5)
cond, y_T, noise = torch.rand(B, 80, generator=g), torch.randn(B, 80, generator=g), torch.randn(T - 2, B, 80, generator=g)
s = shard_rows(B, rank, world)
mine = ddpm.sample(cond[s].to(dev), 500.0, y_init=y_T[s], noise=noise[:, s], stats_group=dist.group.WORLD)
parts = [torch.empty(shard_rows(B, r, world).stop - shard_rows(B, r, world).start, 80, device=dev) for r in range(world)]
dist.all_gather(parts, mine.contiguous()) if len({p.shape[0] for p in parts}) == 1 else [dist.broadcast(parts[r] if r != rank else mine.contiguous(), src=r) for r in range(world)]
parts[rank] = mine
if rank == 0:
    full = ddpm.sample(cond.to(dev), 500.0, y_init=y_T, noise=noise)
    sep = torch.cat([ddpm.sample(cond[shard_rows(B, r, world)].to(dev), 500.0, y_init=y_T[shard_rows(B, r, world)],
                                 noise=noise[:, shard_rows(B, r, world)]) for r in range(world)])
    got = torch.cat(parts)
    rel = lambda a, b: float((a - b).norm() / b.norm())
    print(f"dp_sample_global_stats: world={world} rows={B}: sharded+all-reduced vs un-sharded rel-L2 {rel(got, full):.2e}; "
          f"per-shard statistics vs un-sharded {rel(sep, full):.2e}")
    assert rel(got, full) < 1e-3
dist.barrier()
dist.destroy_process_group()
