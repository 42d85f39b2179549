"""Exponential moving average of model parameters with a fused multi-tensor update.

Same class name, constructor and state_dict layout (`n_averaged`, `module.*`) as the
reference (ddpm_opt/ema.py:3-14, a `torch.optim.swa_utils.AveragedModel` subclass with
`use_buffers=True`), so checkpoints load unchanged.  On CUDA, `update_parameters` is ONE
kernel launch over all tensors (C-ABI `diffsg_ema_update_multi`) instead of a Python loop
of ~3 ops per tensor; on a CPU-resident average (the reference default `device="cpu"`)
the stock AveragedModel arithmetic is what runs, exactly as in the reference.
"""
from __future__ import annotations

import itertools

import torch
from torch.optim.swa_utils import AveragedModel

from . import _lib


class ExponentialMovingAverage(AveragedModel):
    def __init__(self, model, decay, device="cpu"):
        def ema_avg(avg_model_param, model_param, num_averaged):
            return decay * avg_model_param + (1 - decay) * model_param

        super().__init__(model, device, ema_avg, use_buffers=True)
        self.decay = float(decay)
        self._tables = None
        self._n_host = None     # host mirror of n_averaged: the fused path never reads the device counter back

    def _pairs(self, model):
        own = itertools.chain(self.module.parameters(), self.module.buffers())
        src = itertools.chain(model.parameters(), model.buffers())
        return list(zip(own, src))

    def _device_tables(self, pairs):
        key = tuple((a.data_ptr(), b.data_ptr(), a.numel()) for a, b in pairs)
        if self._tables is None or self._tables[0] != key:
            dev = pairs[0][0].device
            avg = torch.tensor([a.data_ptr() for a, _ in pairs], dtype=torch.int64, device=dev)
            src = torch.tensor([b.data_ptr() for _, b in pairs], dtype=torch.int64, device=dev)
            sizes = torch.tensor([a.numel() for a, _ in pairs], dtype=torch.int64, device=dev)
            self._tables = (key, avg, src, sizes, max(a.numel() for a, _ in pairs))
        return self._tables[1:]

    @torch.no_grad()
    def update_parameters(self, model):
        pairs = self._pairs(model)
        fused = bool(pairs) and all(
            a.is_cuda and b.is_cuda and a.device == b.device and a.dtype == torch.float32
            and b.dtype == torch.float32 and a.is_contiguous() and b.is_contiguous() for a, b in pairs)
        if not fused:
            return super().update_parameters(model)
        avg, src, sizes, max_size = self._device_tables(pairs)
        lib = _lib.load()
        # one device -> host read at most (first fused call, or after load_state_dict replaced the counter); afterwards
        # the host mirror decides "first update copies" (ddpm_opt/ema.py:10-14 via AveragedModel) without a stream sync
        if self._n_host is None or self._n_host[0] != (self.n_averaged.data_ptr(), self.n_averaged._version):
            self._n_host = [None, int(self.n_averaged.item())]
        first = self._n_host[1] == 0
        with torch.cuda.device(pairs[0][0].device):
            _lib.check(lib.diffsg_ema_update_multi(avg.data_ptr(), src.data_ptr(), sizes.data_ptr(), len(pairs),
                                                   max_size, self.decay, 1 if first else 0, _lib.stream_ptr()),
                       "diffsg_ema_update_multi")
        self.n_averaged += 1
        self._n_host = [(self.n_averaged.data_ptr(), self.n_averaged._version), self._n_host[1] + 1]
        # the kernel wrote the averaged parameters through raw pointers: tensor versions did not move, so tell the
        # averaged module's kernel plan (if it has one) that its packed weights are stale
        if hasattr(self.module, "mark_params_changed"):
            self.module.mark_params_changed()
