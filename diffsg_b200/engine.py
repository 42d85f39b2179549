"""Kernel plan bound to a `UNet1D` module: owns the C-ABI plan, the packed parameter blob
and the hoisted time tables, and re-packs when the module's parameters change."""
from __future__ import annotations

import contextlib
import ctypes as C

import torch

from . import _lib, tc_packer
from .packer import lower, pack_params, time_table

PRECISIONS = ("auto", "fp32", "fp16x2", "fp16x3")


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.DiffsgError(
                "diffsg_b200 runs on CUDA (sm_100a) only: got a tensor on "
                f"'{t.device}'. There is no CPU implementation; move the module and inputs to a B200.")


@contextlib.contextmanager
def _fp32_matmul():
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


class UNetEngine:
    """precision:
         "fp32"   exact-fp32 warp-row engine (CUDA cores; any topology)
         "fp16x2" tcgen05 engine, activations split into fp16 (hi, lo), weights fp16
         "fp16x3" tcgen05 engine, additionally the fp16 residual of the weights (~fp32 accuracy)
         "auto"   "fp16x2" when the topology fits the tensor-core engine, else "fp32"
    """

    def __init__(self, model, precision="auto"):
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {PRECISIONS}")
        params = list(model.parameters())
        self.device = params[0].device
        if self.device.type != "cuda":
            raise _lib.DiffsgError(f"UNet1D parameters live on '{self.device}'; diffsg_b200 needs a CUDA device")
        self.model = model
        self.lib = _lib.load()
        self.program = lower(model)
        p = self.program
        cfg = _lib.Cfg(abi_version=_lib.ABI_VERSION, input_dim=p.input_dim, cond_dim=p.cond_dim,
                       max_width=p.max_width, n_skip=len(p.skip_widths), skip_floats=sum(p.skip_widths),
                       tt_stride=max(p.tt_stride, 4), tt_rows=0, in_buf=p.in_buf, out_buf=p.out_buf,
                       device=self.device.index if self.device.index is not None else torch.cuda.current_device())
        self._ops = p.op_array()
        sw = (C.c_int32 * max(len(p.skip_widths), 1))(*p.skip_widths)
        handle = C.c_void_p()
        _lib.check(self.lib.diffsg_plan_create(C.byref(cfg), self._ops, len(p.ops), sw, C.byref(handle)),
                   "diffsg_plan_create")
        self.handle = handle
        why = tc_packer.supported(model)
        if precision == "auto":
            precision = "fp32" if why else "fp16x2"
        if precision != "fp32" and why:
            raise _lib.DiffsgError(f"precision {precision!r} needs the tensor-core engine, which does not support: {why}")
        self.precision = precision
        self.tc = None
        if precision != "fp32":
            self.tc = tc_packer.lower_tc(model, nterms=3 if precision == "fp16x3" else 2)
            st, ch, ep = self.tc.arrays()
            self._tc_arrays = (st, ch, ep, (C.c_int32 * max(len(self.tc.skip_widths), 1))(*self.tc.skip_widths))
            prog = _lib.TcProgramC(stages=st.ctypes.data, chunks=ch.ctypes.data, epis=ep.ctypes.data,
                                   skip_widths=C.cast(self._tc_arrays[3], C.c_void_p), n_stages=len(st),
                                   n_chunks=len(ch), n_epi=len(ep), n_skip=len(self.tc.skip_widths),
                                   nterms=self.tc.nterms, tt_stride=max(self.tc.tt_stride, 4))
            _lib.check(self.lib.diffsg_plan_attach_tc(self.handle, C.byref(prog)), "diffsg_plan_attach_tc")
            built = (int(self.lib.diffsg_plan_query(self.handle, 6)), int(self.lib.diffsg_plan_query(self.handle, 7)))
            if built != (tc_packer.CHUNK_K, tc_packer.MAX_W):
                raise _lib.DiffsgError(f"libdiffsg_b200.so was built for (chunk, region) = {built}, the packer lowers for "
                                       f"{(tc_packer.CHUNK_K, tc_packer.MAX_W)}: rebuild (diffsg_b200._lib.build_library(force=True))")
            _lib.check(self.lib.diffsg_plan_set_engine(self.handle, _lib.ENGINE_TC), "diffsg_plan_set_engine")
        self._sig = None
        self._blob = None
        self._tc_blobs = None
        self._tables = {}       # key -> (fp32 time table, fp16 step image table or None)
        self._bound_table = None
        self._stat_ws = None

    def info(self) -> dict:
        q = lambda w: int(self.lib.diffsg_plan_query(self.handle, w))
        return dict(precision=self.precision, engine="tcgen05" if q(0) == _lib.ENGINE_TC else "simt-fp32",
                    tc_ctas_per_sm=q(1), tc_smem_bytes=q(2), tc_grid_max=q(3), simt_warps_per_cta=q(5))

    def __del__(self):
        h = getattr(self, "handle", None)
        if h:
            try:
                self.lib.diffsg_plan_destroy(h)
            except Exception:
                pass
            self.handle = None

    # ------------------------------------------------------------------ parameter binding
    def _signature(self):
        return (self.model._param_epoch,) + tuple((p.data_ptr(), p._version) for p in self.model.parameters())

    def refresh(self):
        """Re-pack the blob if any parameter changed (optimizer step, load_state_dict, ...)."""
        sig = self._signature()
        if sig != self._sig:
            with _fp32_matmul():
                if self.tc is None:
                    self._blob = pack_params(self.program, self.device)
                else:
                    self._tc_blobs = tc_packer.pack_tc_weights(self.tc, self.device)
            self._tables.clear()
            self._bound_table = None
            self._sig = sig

    def _bind(self, table: torch.Tensor, images: torch.Tensor | None = None):
        """Bind the packed weights plus one time table (and, for the tensor-core sampler, its step images)."""
        if self._bound_table is not table:
            if self.tc is None:
                _lib.check(self.lib.diffsg_plan_set_weights(self.handle, self._blob.data_ptr(), self._blob.numel(),
                                                            table.data_ptr(), table.shape[0]),
                           "diffsg_plan_set_weights")
            else:
                hi, lo, params = self._tc_blobs
                _lib.check(self.lib.diffsg_plan_set_tc_weights(
                    self.handle, hi.data_ptr(), lo.data_ptr() if lo is not None else None, hi.numel() * 2,
                    params.data_ptr(), params.numel(), table.data_ptr(), table.shape[0],
                    images.data_ptr() if images is not None else None, images.shape[0] if images is not None else 0,
                    images.shape[1] * 2 if images is not None else 0), "diffsg_plan_set_tc_weights")
            self._bound_table = table
            self._bound_keep = (table, images)

    def _time_table(self, t_values):
        with _fp32_matmul():
            if self.tc is None:
                return time_table(self.model, self.program, t_values)
            return tc_packer.time_table_tc(self.model, self.tc, t_values)

    def step_table(self, T: int):
        """Time-bias table for the sampler: row i <-> t = i / T (reference MSR.py:126).  -> (fp32 table, fp16 step
        images or None): the tensor-core sampler streams the images as bias chunks, forward reads the fp32 rows."""
        key = ("steps", T)
        if key not in self._tables:
            t = torch.arange(T, device=self.device) / T
            table = self._time_table(t)
            self._tables[key] = (table, tc_packer.time_images(self.tc, table) if self.tc is not None else None)
        return self._tables[key]

    def check_status(self, reset=True):
        """Raise if a tensor-core kernel of this plan flagged a raw operand outside the fp16 range since the last
        check (the fp16-split engines saturate it: the call's results are not within tolerance of fp32).
        Synchronises the current stream."""
        if self.tc is None:
            return
        flags = C.c_int32(0)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.diffsg_plan_status(self.handle, C.byref(flags), 1 if reset else 0, _lib.stream_ptr()),
                       "diffsg_plan_status")
        if flags.value & 1:
            raise _lib.DiffsgError(
                f"precision {self.precision!r}: an un-normalised operand (y_t or the residual stream) exceeded the fp16 range "
                "(|x| > 65504); the fp16-split tensor-core engine saturates such values, so this call's results are not "
                "within tolerance of the fp32 reference.  Re-run with model.precision = 'fp32'.")

    def _table_for(self, tv: torch.Tensor):
        """fp32 time table for arbitrary time values `tv` [B] -> (table, row index per element).  The distinct values
        are found with torch.unique (one host sync, needed to size the table); the table itself is cached per
        distinct-value set, so repeated calls on the same time grid (a sampling loop, a training schedule) reuse it."""
        uniq, inv = torch.unique(tv, return_inverse=True)
        if uniq.numel() <= 4096:
            key = ("values", tuple(uniq.tolist()))
            if key not in self._tables:
                if sum(1 for k in self._tables if k[0] == "values") >= 64:      # bound the cache
                    for k in [k for k in self._tables if k[0] == "values"]:
                        if self._tables[k][0] is not self._bound_table:
                            del self._tables[k]
                self._tables[key] = (self._time_table(uniq), None)
            return self._tables[key][0], inv
        return self._time_table(uniq), inv

    # ------------------------------------------------------------------ forward
    def forward(self, x, t, cond, cond_mask, t_index=None, n_steps=None):
        """`t` [1, B] arbitrary time values (the distinct ones are found with torch.unique: one host sync), or
        `t_index` [B] integer steps of an `n_steps` schedule (cached step table, no sync)."""
        _require_cuda(x, t, cond, cond_mask, t_index)
        with torch.cuda.device(self.device):        # streams, allocations and the C side all follow the model's device
            return self._forward(x, t, cond, cond_mask, t_index, n_steps)

    def _forward(self, x, t, cond, cond_mask, t_index, n_steps):
        self.refresh()
        x2 = _f32c(x).reshape(-1, self.program.input_dim)
        B = x2.shape[0]
        if B == 0:      # empty batch: nothing to launch (the reference returns an empty tensor too)
            return torch.empty(0, self.program.input_dim, dtype=torch.float32, device=self.device)
        cond2 = _f32c(cond).reshape(B, self.program.cond_dim)
        images = None
        if t_index is not None:
            table, images = self.step_table(int(n_steps))
            inv = t_index.detach().reshape(-1)
            if inv.numel() == 1 and B > 1:
                inv = inv.expand(B)
        else:
            tv = t.detach().reshape(-1).to(torch.float32)
            if tv.numel() == 1 and B > 1:
                tv = tv.expand(B)
            table, inv = self._table_for(tv)
        self._bind(table, images)
        mask = None
        if cond_mask is not None:
            mask = _f32c(cond_mask).reshape(-1)
            if mask.numel() == 1 and B > 1:
                mask = mask.expand(B).contiguous()
        t_idx = inv.to(torch.int32).contiguous()
        eps = torch.empty(B, self.program.input_dim, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.diffsg_unet_forward(self.handle, x2.data_ptr(), t_idx.data_ptr(), cond2.data_ptr(),
                                                mask.data_ptr() if mask is not None else None, eps.data_ptr(),
                                                B, _lib.stream_ptr()), "diffsg_unet_forward")
        return eps

    # ------------------------------------------------------------------ sampler
    def sample(self, cond, y_init, coef, T, omega, *, noise=None, seed=0, offset=0, norm_steps=4,
               rec_y=None, rec_eps=None, stats_reduce=None):
        """In-place reverse diffusion on `y_init` ([B, M] fp32 CUDA); returns it.

        `stats_reduce(t)`: optional callable that sums the fp64 tensor t = [sum y, sum y^2, count] in place over
        all shards of ONE logical batch (e.g. `dist.all_reduce`).  The re-normalised steps then use whole-batch
        statistics — the reference's semantics for the un-sharded call — at the cost of one tiny collective per
        re-normalised step; without it each shard normalises over its own rows."""
        _require_cuda(cond, y_init, noise, rec_y, rec_eps)
        if y_init.shape[0] == 0 and stats_reduce is None:
            return y_init
        with torch.cuda.device(self.device):
            return self._sample(cond, y_init, coef, T, omega, noise, seed, offset, norm_steps, rec_y, rec_eps, stats_reduce)

    def _sample(self, cond, y_init, coef, T, omega, noise, seed, offset, norm_steps, rec_y, rec_eps, stats_reduce):
        self.refresh()
        table, images = self.step_table(T)
        self._bind(table, images)
        B = y_init.shape[0]
        M = self.program.input_dim
        if self._stat_ws is None or self._stat_ws.numel() < 2 * T:
            self._stat_ws = torch.zeros(2 * max(T, 64), dtype=torch.float64, device=self.device)
        coef_arr = (C.c_float * (3 * T))(*[float(v) for v in coef])
        args = _lib.SampleArgs(cond_dev=cond.data_ptr(), y_dev=y_init.data_ptr(),
                               noise_dev=noise.data_ptr() if noise is not None else None,
                               rec_y_dev=rec_y.data_ptr() if rec_y is not None else None,
                               rec_eps_dev=rec_eps.data_ptr() if rec_eps is not None else None,
                               stat_ws_dev=self._stat_ws.data_ptr(),
                               coef_host=C.cast(coef_arr, C.c_void_p), B=B, T=T, norm_steps=norm_steps,
                               omega=float(omega), pad_=0, philox_seed=int(seed) & (2**64 - 1),
                               philox_offset=int(offset) & (2**64 - 1))
        if stats_reduce is None:
            _lib.check(self.lib.diffsg_sample(self.handle, C.byref(args), _lib.stream_ptr()), "diffsg_sample")
            return y_init
        # whole-batch statistics over several shards: step-wise through the same kernels
        n_norm = min(norm_steps, T)
        cnt = torch.tensor([0.0, 0.0, float(B * M)], dtype=torch.float64, device=self.device)
        stats_reduce(cnt)
        n_stat = int(round(float(cnt[2])))                  # element count of the logical batch (the one host sync)
        for i in range(T - 1, T - 1 - n_norm, -1):
            tot = torch.zeros(3, dtype=torch.float64, device=self.device)
            if B > 0:
                _lib.check(self.lib.diffsg_sample_steps(self.handle, C.byref(args), i, i, 0, _lib.stream_ptr()), "diffsg_sample_steps")
                tot[:2] = self._stat_ws[2 * i:2 * i + 2]
            stats_reduce(tot)
            if B > 0:
                plane = rec_y[T - 1 - i] if rec_y is not None else None
                _lib.check(self.lib.diffsg_sample_renorm(y_init.data_ptr(), plane.data_ptr() if plane is not None else None,
                                                         tot.data_ptr(), B * M, n_stat, _lib.stream_ptr()), "diffsg_sample_renorm")
        if B > 0 and T - 1 - n_norm >= 0:
            _lib.check(self.lib.diffsg_sample_steps(self.handle, C.byref(args), T - 1 - n_norm, 0, 1, _lib.stream_ptr()), "diffsg_sample_steps")
        return y_init


def _sample_args(engine, cond, y, coef_arr, T, omega, noise=None, seed=0, offset=0, norm_steps=4, rec_y=None, rec_eps=None):
    return _lib.SampleArgs(cond_dev=cond.data_ptr(), y_dev=y.data_ptr(),
                           noise_dev=noise.data_ptr() if noise is not None else None,
                           rec_y_dev=rec_y.data_ptr() if rec_y is not None else None,
                           rec_eps_dev=rec_eps.data_ptr() if rec_eps is not None else None,
                           stat_ws_dev=engine._stat_ws.data_ptr(), coef_host=C.cast(coef_arr, C.c_void_p),
                           B=y.shape[0], T=T, norm_steps=norm_steps, omega=float(omega), pad_=0,
                           philox_seed=int(seed) & (2**64 - 1), philox_offset=int(offset) & (2**64 - 1))


def sampler_pass_eps(engine: UNetEngine, cond, y_t, step: int, coef, T: int):
    """(eps_0, eps_1): the unconditional and the conditional prediction of the SAMPLER kernels (the fused path, not
    `forward`) on the state `y_t` [B, M] of reverse step `step`: two single-step launches whose guidance weight turns
    the recorded mixed eps = (1 + omega) eps_1 - omega eps_0 into one pass each (omega = -1 -> eps_0, omega = 0 ->
    eps_1).  Teacher-forced parity checks and bench.py's `parity` record read per-pass errors through this."""
    _require_cuda(cond, y_t)
    B, M = y_t.shape
    with torch.cuda.device(engine.device):
        engine.refresh()
        table, images = engine.step_table(T)
        engine._bind(table, images)
        if engine._stat_ws is None or engine._stat_ws.numel() < 2 * T:
            engine._stat_ws = torch.zeros(2 * max(T, 64), dtype=torch.float64, device=engine.device)
        coef_arr = (C.c_float * (3 * T))(*[float(v) for v in coef])
        rec = torch.empty(T, B, M, dtype=torch.float32, device=engine.device)     # only plane T-1-step is written
        out = []
        for omega in (-1.0, 0.0):
            y = y_t.detach().to(torch.float32).clone().contiguous()
            args = _sample_args(engine, cond, y, coef_arr, T, omega, rec_eps=rec)
            _lib.check(engine.lib.diffsg_sample_steps(engine.handle, C.byref(args), step, step, 0, _lib.stream_ptr()),
                       "diffsg_sample_steps")
            out.append(rec[T - 1 - step].clone())
    return out[0], out[1]


def unet_forward(model, x, t, cond, cond_mask):
    """`UNet1D.forward` entry: inference through the fused kernels; training through
    `diffsg_b200.train` when gradients are required."""
    needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in model.parameters())
    if needs_grad:
        from .train import unet_forward_train
        return unet_forward_train(model, x, t, cond, cond_mask)
    return model.engine().forward(x, t, cond, cond_mask)


def philox_normal(B: int, M: int, step: int, seed: int, offset: int, device) -> torch.Tensor:
    """The sampler's own noise stream for (seed, offset, step) as a [B, M] tensor."""
    out = torch.empty(B, M, dtype=torch.float32, device=device)
    if B == 0:
        return out
    lib = _lib.load()
    with torch.cuda.device(out.device):
        _lib.check(lib.diffsg_philox_normal(out.data_ptr(), B, M, step, int(seed) & (2**64 - 1),
                                            int(offset) & (2**64 - 1), _lib.stream_ptr()), "diffsg_philox_normal")
    return out
