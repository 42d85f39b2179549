"""Script-level entry points of the reference's three front-ends, shared implementation.

The reference scripts each carry a copy of the same training loop and the same evaluation driver
(ddpm_opt/classifier_free_MSR.py:187-236, 248-298; _NU.py:213-264, 306-361; _CO.py:203-252, 293-356).
`fit` is that loop (DataLoader(bs 512, shuffle) -> loss = DDPM(y, x) -> backward -> Adam -> MultiStepLR -> gated EMA,
the same per-epoch print incl. its sum-of-batch-means / sample-count quirk) on this package's trainer: flat
parameter / gradient buffers, one all-reduce when a process group is up, the library's fused Adam (+ EMA) kernel.
`diffsg_b200.{msr,nu,co}` wrap it as `train_ddpm_*` / `load_test_*` with the reference's names, constants and
printed report; every hard-coded constant of the reference is a keyword default here.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .parallel import DataParallelTrainer, MultiStepLR
from .schedule import generate_cosine_schedule, init_weights
from .unet import UNet1D, infer_config_from_state_dict


def default_device():
    """The reference picks cuda:0 when present (classifier_free_MSR.py:205-207); this package has no CPU path."""
    if not torch.cuda.is_available():
        raise _lib.DiffsgError("diffsg_b200 needs a CUDA device (sm_100a); there is no CPU implementation")
    print(f"Found cuda. Device count: {torch.cuda.device_count()}, the 0 is {torch.cuda.get_device_name(0)}")
    return torch.device("cuda:0")


def fit(diffusion_model, X_train, Y_train, *, epochs, lr, milestones, batch_size=512, use_ema=False, warmup_epoch=5,
        cuda_graph=True, verbose=True, seed=None):
    """The reference training loop (classifier_free_MSR.py:213-236) on `diffusion_model`; returns the per-epoch
    losses as the reference prints them.  Shuffling uses torch's CPU generator like `DataLoader(shuffle=True)`
    (seeded by `seed` when given); the last short batch of an epoch is kept, as the reference keeps it."""
    dev = diffusion_model.betas.device
    X = torch.as_tensor(np.asarray(X_train), dtype=torch.float32).to(dev)
    Y = torch.as_tensor(np.asarray(Y_train), dtype=torch.float32).to(dev)
    n = X.shape[0]
    trainer = DataParallelTrainer(diffusion_model, lr=lr, cuda_graph=cuda_graph)
    sched = MultiStepLR(trainer.opt, milestones)
    gen = torch.Generator()
    if seed is not None:
        gen.manual_seed(seed)
    history = []
    for epoch in range(epochs):
        # EMA gate of the reference: `use_ema and epoch > warmup_epoch and cnt > ema_start and cnt % rate == 0`;
        # the step-count part lives in the trainer, the epoch part here
        trainer.use_ema = bool(use_ema) and epoch > warmup_epoch
        perm = torch.randperm(n, generator=gen).to(dev)
        losses = []
        for i in range(0, n, batch_size):
            idx = perm[i:i + batch_size]
            losses.append(trainer.step(Y[idx], X[idx]))
        epoch_loss = float(torch.stack(losses).sum())            # one host read per epoch instead of one per batch
        history.append(epoch_loss / n)
        if verbose:
            print(f"Epoch: {epoch}, Loss: {epoch_loss / n}")
        sched.step()
    return history


def build_ddpm(ddpm_cls, ctor_args, unet_cfg, T, device):
    """UNet1D(**unet_cfg) + DDPM(T, model, *ctor_args, alphas, device, ...) with the scripts' fixed trailing arguments."""
    model = UNet1D(**unet_cfg)
    alphas = 1.0 - generate_cosine_schedule(T)
    return ddpm_cls(T, model, *ctor_args, alphas, device, (1, unet_cfg["input_dim"]), None, 0.1, 0.9999, 10, 5, False)


def train(ddpm_cls, ctor_args, unet_cfg, custom_config, X_train, Y_train, *, T=20, epochs=200, lr, milestones,
          use_ema=False, device=None, **fit_kw):
    device = device or default_device()
    ddpm = build_ddpm(ddpm_cls, ctor_args, unet_cfg, T, device)
    ddpm.custom_config = custom_config
    ddpm.apply(init_weights)
    ddpm.to(device)
    fit(ddpm, X_train, Y_train, epochs=epochs, lr=lr, milestones=milestones, use_ema=use_ema, **fit_kw)
    return ddpm


def load(ddpm_cls, ctor_args, unet_cfg, custom_config, ckpt, *, T=20, device=None):
    """Strict-load a checkpoint (path or state_dict) into a freshly built DDPM, as `load_test_*` does
    (classifier_free_MSR.py:262-271).  `unet_cfg` = the script's hard-coded topology; when the checkpoint was trained
    with another one (the 80-channel hyper-parameters appear nowhere in the reference) the topology is
    inferred from the tensor shapes instead."""
    device = device or default_device()
    sd = ckpt if isinstance(ckpt, dict) else torch.load(ckpt, map_location="cpu")
    inferred = infer_config_from_state_dict(sd)
    if any(inferred[k] != (tuple(v) if isinstance(v, (list, tuple)) else v) for k, v in unet_cfg.items() if k in inferred):
        unet_cfg = inferred
    ddpm = build_ddpm(ddpm_cls, ctor_args, unet_cfg, T, device)
    ddpm.custom_config = custom_config
    ddpm.load_state_dict(sd)
    return ddpm.to(device)


def report(title_rows, summary_rows, precision=4):
    """The tensor dumps + summary lines `load_test_*` prints (first 20 rows of each, classifier_free_MSR.py:289-298)."""
    torch.set_printoptions(precision=precision, sci_mode=False)
    np.set_printoptions(precision=precision, suppress=True)
    for name, t in title_rows:
        print(f"{name}:\n", t[:20])
    for line in summary_rows:
        print(line)
