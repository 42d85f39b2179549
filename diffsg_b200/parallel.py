"""Data-parallel plumbing: one process per GPU, `torch.distributed` (NCCL over NVLink / NVSwitch
on the GPU box, gloo in CPU tests).

* Sampling shards by instance with NO collective: `shard_rows` gives each rank a contiguous slice
  (batch statistics of the four re-normalised steps are per shard, exactly the reference's
  per-call semantics).
* Training: every parameter and every gradient is a view into one flat fp32 buffer, so the
  gradient exchange is ONE all-reduce of `n_params` floats (6.6 MB for 80c) instead of ~490, and the
  optimiser + EMA are ONE launch of the library's own kernel (`diffsg_adam_step`: Adam and, on the steps the
  reference's gate selects, the EMA of the freshly updated parameters in the same pass).
  (The reference has no distributed code at all: SURVEY §2a.)
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib


def shard_rows(n_rows: int, rank: int, world: int) -> slice:
    """Contiguous row shard of rank `rank`: sizes differ by at most one, concatenation == range(n)."""
    base, extra = divmod(n_rows, world)
    start = rank * base + min(rank, extra)
    return slice(start, start + base + (1 if rank < extra else 0))


class FlatParams:
    """Re-homes a module's parameters (and their .grad) into two flat fp32 buffers."""

    def __init__(self, module: torch.nn.Module):
        params = [p for p in module.parameters() if p.requires_grad]
        if not params:
            raise ValueError("module has no trainable parameters")
        dev = params[0].device
        n = sum(p.numel() for p in params)
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in params:
                k = p.numel()
                self.flat[off:off + k].copy_(p.detach().reshape(-1))
                p.data = self.flat[off:off + k].view_as(p)
                p.grad = self.grad[off:off + k].view_as(p)
                off += k
        self.params = params
        self.numel = n

    def zero_grad(self):
        self.grad.zero_()          # keeps the views: never set p.grad = None

    def allreduce_grads(self, group=None, average=True):
        """One collective for the whole model; no-op outside a process group."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=group)
            if average:
                self.grad.div_(dist.get_world_size(group))


class FusedAdam:
    """Adam over ONE flat fp32 buffer through `diffsg_adam_step` (torch.optim.Adam semantics: lr, betas, eps; no
    weight decay, no amsgrad -- what the reference loop constructs, classifier_free_MSR.py:213).  Hyper-parameters
    and the step counter live on the device, so `step()` can be captured in a CUDA graph while `set_lr` (MultiStepLR)
    and the EMA gate keep changing between replays."""

    def __init__(self, flat: torch.Tensor, grad: torch.Tensor, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        if not flat.is_cuda:
            raise _lib.DiffsgError("FusedAdam runs on CUDA only (no CPU implementation)")
        self.flat, self.grad = flat, grad
        self.exp_avg = torch.zeros_like(flat)
        self.exp_avg_sq = torch.zeros_like(flat)
        self.hyper = torch.tensor([lr, betas[0], betas[1], eps, 0.0, 0.0], dtype=torch.float32, device=flat.device)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=flat.device)
        self.lr = float(lr)
        self.lib = _lib.load()

    def set_lr(self, lr: float):
        if float(lr) != self.lr:
            self.lr = float(lr)
            self.hyper[0:1].fill_(self.lr)

    def set_ema(self, decay: float, mode: int):
        """mode 0: no EMA on the next step; 1: the average becomes a copy; 2: avg = decay avg + (1 - decay) p."""
        self.hyper[4:6].copy_(torch.tensor([decay, float(mode)], dtype=torch.float32), non_blocking=True)

    def reset_state(self):
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        self.step_dev.zero_()

    def step(self, ema_flat: torch.Tensor | None = None):
        with torch.cuda.device(self.flat.device):
            _lib.check(self.lib.diffsg_adam_step(self.flat.data_ptr(), self.grad.data_ptr(), self.exp_avg.data_ptr(),
                                                 self.exp_avg_sq.data_ptr(),
                                                 ema_flat.data_ptr() if ema_flat is not None else None, self.flat.numel(),
                                                 self.hyper.data_ptr(), self.step_dev.data_ptr(), _lib.stream_ptr()),
                       "diffsg_adam_step")


class MultiStepLR:
    """lr(epoch) = base * gamma ** #(milestones <= epoch): torch.optim.lr_scheduler.MultiStepLR as the reference
    drives it (one `.step()` per epoch, classifier_free_MSR.py:214,234), for a `FusedAdam`."""

    def __init__(self, optimizer: "FusedAdam", milestones, gamma=0.1):
        self.opt, self.milestones, self.gamma = optimizer, sorted(milestones), gamma
        self.base_lr, self.epoch = optimizer.lr, 0

    def get_last_lr(self):
        return [self.base_lr * self.gamma ** sum(1 for m in self.milestones if m <= self.epoch)]

    def step(self):
        self.epoch += 1
        self.opt.set_lr(self.get_last_lr()[0])


class DataParallelTrainer:
    """eps-MSE training step of `DDPM.forward` with flat-buffer gradient all-reduce, fused Adam and
    fused EMA (reference loop: ddpm_opt/classifier_free_MSR.py:220-232).

    `cuda_graph=True` captures the step once per batch shape and replays it: the ~280 kernels of one
    forward + backward (this library's tcgen05 forward / backward nodes, `diffsg_b200.train`, the elementwise
    glue and the RNG draws of `DDPM.forward`) become ONE graph launch, the fused Adam a second one; the NCCL
    all-reduce of the flat gradient stays an ordinary stream-ordered call between the two.  wgrad accumulates
    straight into the flat gradient buffer (the parameters' `.grad` are views of it).  DESIGN.md §5.4."""

    def __init__(self, ddpm, lr=0.005, ema_device_update=True, cuda_graph=False):
        self.ddpm = ddpm
        self.flat = FlatParams(ddpm.model)
        self.cuda_graph = bool(cuda_graph)
        self.opt = FusedAdam(self.flat.flat, self.flat.grad, lr=lr)      # the library's own Adam (+ EMA) kernel; CUDA only
        self.step_count = 0
        self.use_ema = False
        self._ema_flat = None      # EMA parameters re-homed into one flat buffer (same order as self.flat)
        self._graphs = {}          # (y shape, cond shape) -> (fwd/bwd graph, optimiser graph, static y, static cond, static loss)

    def set_lr(self, lr):
        self.opt.set_lr(lr)

    def _ema_buffer(self):
        """Flat view of the EMA copy's parameters (the averaged module keeps working: its tensors become views)."""
        if self._ema_flat is None:
            own = [p for p in self.ddpm.ema.module.parameters()]
            if len(own) != len(self.flat.params) or any(a.shape != b.shape for a, b in zip(own, self.flat.params)):
                raise _lib.DiffsgError("EMA module and model disagree on parameter shapes")
            dev = self.flat.flat.device
            buf = torch.empty(self.flat.numel, dtype=torch.float32, device=dev)
            off = 0
            with torch.no_grad():
                for p in own:
                    k = p.numel()
                    buf[off:off + k].copy_(p.detach().reshape(-1).to(dev))
                    p.data = buf[off:off + k].view_as(p)
                    off += k
            self._ema_flat = buf
        return self._ema_flat

    def _fwd_bwd(self, y, cond):
        self.flat.zero_grad()
        loss = self.ddpm(y, cond)
        loss.backward()
        return loss.detach()

    def _capture(self, y, cond):
        """Warm up on a side stream (lazy library / optimiser state) WITHOUT changing the model: lr = 0 during
        the warm-up steps, Adam's moments and step counter reset afterwards; then capture the two graphs."""
        sy, sc = y.clone(), cond.clone()
        lr = self.opt.lr
        side = torch.cuda.Stream(device=sy.device)
        side.wait_stream(torch.cuda.current_stream(sy.device))
        with torch.cuda.stream(side):
            self.opt.set_lr(0.0)
            for _ in range(3):
                self._fwd_bwd(sy, sc)
                self.opt.step()
            self.opt.set_lr(lr)
            self.opt.reset_state()
        torch.cuda.current_stream(sy.device).wait_stream(side)
        g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(g1):
            sloss = self._fwd_bwd(sy, sc)
        with torch.cuda.graph(g2):
            self.opt.step(self._ema_buffer() if self.use_ema else None)
        return g1, g2, sy, sc, sloss

    def step(self, y, cond):
        d = self.ddpm
        # EMA gate of the reference loop (classifier_free_MSR.py:227-229), decided on the host before the launch
        ema_now = self.use_ema and (self.step_count + 1) > d.ema_start and (self.step_count + 1) % d.ema_update_rate == 0
        ema_flat = None
        if self.use_ema:
            ema_flat = self._ema_buffer()
            first = getattr(self, "_ema_updates", 0) == 0
            self.opt.set_ema(d.ema.decay, 0 if not ema_now else (1 if first else 2))
        if self.cuda_graph:
            key = (tuple(y.shape), tuple(cond.shape), self.use_ema)
            if key not in self._graphs:
                self._graphs[key] = self._capture(y, cond)
            g1, g2, sy, sc, sloss = self._graphs[key]
            sy.copy_(y)
            sc.copy_(cond)
            g1.replay()
            self.flat.allreduce_grads()
            g2.replay()
            loss = sloss.clone()
        else:
            loss = self._fwd_bwd(y, cond)
            self.flat.allreduce_grads()
            self.opt.step(ema_flat)
        self.ddpm.model.mark_params_changed()      # the parameters are views of the flat buffer Adam just updated
        self.step_count += 1
        if ema_now:                                # done inside the Adam launch; keep the module's bookkeeping in step
            self._ema_updates = getattr(self, "_ema_updates", 0) + 1
            d.ema.n_averaged += 1
            if hasattr(d.ema.module, "mark_params_changed"):
                d.ema.module.mark_params_changed()
        return loss
