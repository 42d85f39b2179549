#!/usr/bin/env python
"""Headline benchmark: CFG-DDPM solutions/s on the 80-channel MSR configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--mode sample|train]

One "step" = one full pass of the hot path over one batch: `DDPM.sample` for `--rows`
synthetic condition rows (default 1 Mi per GPU; BASELINE.json configs[1]) = T=20 reverse
steps x 2 UNet passes + guidance + posterior update + 4 batch re-normalisations.
Rank 0 prints ONE JSON line (see the task contract):
  value         device-resident throughput (inputs in HBM when the timed region starts), weak scaling
  e2e           the same through the public API with pinned-host inputs/outputs copied inside the timed region
  roofline      algorithmic FLOP/s of the sampler kernel against the measured bf16 tensor peak (+ measured DRAM traffic)
  strong        BASELINE's own configuration: ONE 1 Mi-row batch split over the N GPUs
  precision_modes  the ~fp32 sibling mode (fp16x3) on the same workload
  parity        the bench network / noise / omega checked against the oracle in this run (per-pass eps, objective)
  cpu_baseline  the unmodified reference (baseline/_ref, torch CPU, all host threads) on a bounded sample, same run
  gpu_eager_baseline  the unmodified reference on cuda:0 (stock PyTorch eager): the like-for-like kernel to beat

`--impl reference` times the reference's own implementation of the path (the UNMODIFIED reference from
baseline/_ref when it is there, else the oracle port of oracle/ddpm_oracle.py, bit-identical to it under
make_golden.py) on the box's host cores; under torchrun only rank 0 runs it.
`--mode train` measures BASELINE configs[4] instead: the eps-MSE training step (+ gated EMA) of the 80c net on
synthetic data, data-parallel with one flat-gradient NCCL all-reduce per step.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
REF_DIR = ROOT / "baseline" / "_ref"

T = 20
OMEGA = 500.0
NET = dict(input_dim=80, proj_dim=128, cond_dim=80, dims=(64, 32, 16, 8), is_attn=(False,) * 4,
           middle_attn=False, n_blocks=2)   # 80c = the 3c script with M=80 (ASSUMED)
METRIC = "CFG-DDPM solutions/sec (80c MSR, T=20, omega=500)"
TRAIN_METRIC = "eps-MSE training samples/sec (80c MSR, Adam + EMA, data-parallel)"
WORKLOAD = ("BASELINE configs[1]: MSR 80c CFG-DDPM sampling (assumed UNet1D 80/128/80/(64,32,16,8)/2, init_weights N(0,0.01) "
            "seed 0), synthetic rand(B,80) conditions, rows sharded per GPU, no collective")
FALLBACK_PEAKS = {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0}
TRAFFIC_FILE = ROOT / "profiles" / "r02_traffic_tc.json"


def build_model(device):
    import diffsg_b200 as D
    torch.manual_seed(0)
    model = D.UNet1D(**NET)          # no 80c checkpoint exists: reference init, seed 0
    alphas = 1.0 - D.generate_cosine_schedule(T)
    ddpm = D.msr.DDPM(T, model, NET["input_dim"], 20.0, alphas, device, (1, NET["input_dim"]),
                      {"scaler_min": 0.5, "scaler_max": 2.5, "W": 20.0}, 0.1, 0.9999, 10, 5, False)
    ddpm.apply(D.init_weights)       # the reference's own init (diffusion.py:82-84); finite at omega=500
    return ddpm.to(device)


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return json.loads(p.read_text()), "measured"
    return dict(FALLBACK_PEAKS), "fallback"


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        return os.cpu_count() or 1


class ClockSampler:
    """Samples SM clocks + throttle reasons via NVML every 200 ms while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.reasons.update(n for bit, n in names.items() if mask & bit)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        if self.nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------ the reference itself
def reference_ddpm(sd, device):
    """The UNMODIFIED reference classes from baseline/_ref (a plain copy of the reference tree made by
    __graft_entry__.build()), carrying the bench network's weights.  None when the copy is absent."""
    if not (REF_DIR / "ddpm_opt" / "classifier_free_MSR.py").exists():
        return None
    saved = {k: v for k, v in sys.modules.items() if k == "ddpm_opt" or k.startswith("ddpm_opt.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, str(REF_DIR))
    try:
        from ddpm_opt.UNetCF import UNet1D as RefUNet
        from ddpm_opt.classifier_free_MSR import DDPM as RefDDPM
        from ddpm_opt.diffusion import generate_cosine_schedule as ref_schedule
        model = RefUNet(**NET)
        ddpm = RefDDPM(T, model, NET["input_dim"], 20.0, 1.0 - ref_schedule(T), device, (1, NET["input_dim"]),
                       {"scaler_min": 0.5, "scaler_max": 2.5, "W": 20.0}, 0.1, 0.9999, 10, 5, False)
        ddpm.load_state_dict(sd)
        return ddpm.to(device)
    finally:
        sys.path.remove(str(REF_DIR))
        for k in [k for k in sys.modules if k == "ddpm_opt" or k.startswith("ddpm_opt.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def cpu_reference_run(sd, rows, reps, warmup):
    """The reference algorithm on the host cores -> (times, kind).  kind "reference": the unmodified `DDPM.sample`
    of baseline/_ref (its own CPU noise draws); "port": the oracle port (the one sanctioned use of oracle/ here)."""
    g = torch.Generator().manual_seed(0)
    cond = torch.rand(rows, NET["cond_dim"], generator=g)
    ref = reference_ddpm(sd, "cpu")
    times = []
    with torch.no_grad():
        for r in range(warmup + reps):
            if ref is not None:
                torch.manual_seed(1000 + r)
                t0 = time.perf_counter()
                ref.sample(cond, OMEGA)
            else:
                from oracle import ddpm_oracle as O
                y_T, steps = O.draw_noise(rows, (1, NET["input_dim"]), T, 1000 + r)
                t0 = time.perf_counter()
                O.sample(sd, T, cond, OMEGA, y_T, steps)
            if r >= warmup:
                times.append(time.perf_counter() - t0)
    return times, ("reference" if ref is not None else "port")


def gpu_eager_run(sd, dev, rows, reps):
    """The unmodified reference on the GPU through its stock code path (PyTorch eager; noise drawn on the CPU and
    copied H2D every step, classifier_free_MSR.py:115,129) -> solutions/s, or None without baseline/_ref."""
    ref = reference_ddpm(sd, dev)
    if ref is None:
        return None
    g = torch.Generator().manual_seed(0)
    cond = torch.rand(rows, NET["cond_dim"], generator=g).to(dev)
    times = []
    with torch.no_grad():
        for r in range(1 + reps):
            torch.manual_seed(2000 + r)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ref.sample(cond, OMEGA)
            torch.cuda.synchronize()
            if r >= 1:
                times.append(time.perf_counter() - t0)
    return rows / (sum(times) / len(times))


def sample_config(rows_per_gpu, stats="per shard (reference per-call semantics)"):
    """The workload definition shared verbatim by both arms (how each arm executes it is reported outside `config`)."""
    return {"workload": WORKLOAD, "rows_per_gpu": rows_per_gpu, "T": T, "omega": OMEGA, "noise": "in-kernel Philox4x32-10",
            "batch_stats": stats,
            "l2": f"inputs+state per step {2 * rows_per_gpu * NET['input_dim'] * 4 / 2**20:.0f} MiB > 126 MB L2; no explicit flush"}


def run_reference(args, rank):
    if rank != 0:
        return
    torch.set_num_threads(host_threads())     # torchrun exports OMP_NUM_THREADS=1; the reference arm is entitled to every core
    ddpm = build_model("cpu")
    sd = {k: v.detach().clone() for k, v in ddpm.state_dict().items()}
    rows = args.ref_rows
    times, kind = cpu_reference_run(sd, rows, args.steps, args.warmup)
    dt = sum(times) / len(times)
    val = rows / dt
    cores = torch.get_num_threads()
    what = ("unmodified reference DDPM.sample (baseline/_ref)" if kind == "reference"
            else "reference algorithm (oracle port, bit-identical to reference DDPM.sample)")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "solutions/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": sample_config(args.rows),
            "cpu_baseline": {"value": val, "unit": "solutions/s", "cores": cores, "kind": kind,
                             "sample": f"{rows} rows per step x {args.steps} steps of the same workload (rows are independent; the "
                                       f"reference itself samples 512-row loader batches), torch CPU, {what}"},
            "e2e": {"value": val, "unit": "solutions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ parity in the bench run
def parity_record(ddpm, dev, rows=4096):
    """The benchmark's own configuration (80c init_weights net, Philox noise, omega = 500) against the oracle:
    the GPU and the oracle are fed the SAME Philox planes; free-running trajectories are chaotic at omega = 500
    (SURVEY H1), so the gates are (a) per-pass eps of the SAMPLER kernels, teacher-forced on the GPU's trajectory,
    rel-L2 <= 1e-3, and (b) the MSR objective of the two free-running results within 0.5 %."""
    from oracle import ddpm_oracle as O       # checker only
    from diffsg_b200 import objectives
    from diffsg_b200.engine import sampler_pass_eps
    M, Cd = NET["input_dim"], NET["cond_dim"]
    seed, offset = 77, 12345
    g = torch.Generator().manual_seed(5)
    cond = torch.rand(rows, Cd, generator=g)
    sd = {k: v.detach().cpu().clone() for k, v in ddpm.state_dict().items()}
    y_T = torch.as_tensor(O.philox_normal(rows, M, T, seed, offset), dtype=torch.float32)
    planes = [torch.as_tensor(O.philox_normal(rows, M, i, seed, offset), dtype=torch.float32) for i in range(T - 1, 1, -1)]
    t0 = time.perf_counter()
    with torch.no_grad():
        y_ref = O.sample(sd, T, cond, OMEGA, y_T.reshape(rows, 1, M), [p.reshape(rows, 1, M) for p in planes]).reshape(rows, M)
    oracle_s = time.perf_counter() - t0
    eng = ddpm.model.engine()
    cond_d = cond.to(dev)
    rec_y = torch.empty(T, rows, M, device=dev)
    y = y_T.to(dev).clone()
    coef = ddpm.step_coefficients()
    eng.sample(cond_d, y, coef, T, OMEGA, seed=seed, offset=offset, rec_y=rec_y)      # in-kernel Philox: the same planes
    eng.check_status()
    gains = cond * 2.0 + 0.5
    obj_gpu = float(objectives.msr_decode_rate(y, gains.to(dev), 20.0).mean())
    obj_ref = float(O.msr_rate(20.0 * O.msr_decode(y_ref), gains).mean())
    states = [y_T.to(dev)] + [rec_y[j] for j in range(T - 1)]
    worst, worst_mix = 0.0, 0.0
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
    with torch.no_grad():
        for j, i in enumerate(range(T - 1, -1, -1)):
            e0, e1 = sampler_pass_eps(eng, cond_d, states[j], i, coef, T)
            yt = states[j].cpu()
            tt = torch.full((1, rows), i, dtype=torch.float32) / T
            w0 = O.unet_forward(sd, yt, tt, cond, torch.zeros(rows, 1))
            w1 = O.unet_forward(sd, yt, tt, cond, torch.ones(rows, 1))
            worst = max(worst, rel(e0.cpu(), w0), rel(e1.cpu(), w1))
            worst_mix = max(worst_mix, rel((1 + OMEGA) * e1.cpu().double() - OMEGA * e0.cpu().double(),
                                           (1 + OMEGA) * w1.double() - OMEGA * w0.double()))
    return {"rows": rows, "noise": "identical Philox planes on both sides", "engine": eng.precision,
            "worst_per_pass_eps_rel_l2": worst, "gate_per_pass": 1e-3, "mixed_eps_rel_l2_omega500": worst_mix,
            "objective_mean_gpu": obj_gpu, "objective_mean_oracle": obj_ref, "objective_ratio": obj_gpu / obj_ref,
            "gate_objective": 5e-3, "pass": bool(worst <= 1e-3 and abs(obj_gpu / obj_ref - 1) <= 5e-3),
            "oracle": "oracle/ddpm_oracle.py (bit-identical to the reference, oracle/make_golden.py)", "oracle_seconds": oracle_s}


# ------------------------------------------------------------------------------------------ training mode
def run_train(args, rank, local_rank, world):
    import torch.distributed as dist
    from diffsg_b200 import _lib
    from diffsg_b200.parallel import DataParallelTrainer
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ddpm = build_model(dev)
    tr = DataParallelTrainer(ddpm, lr=1e-3, cuda_graph=not args.no_graph)
    tr.use_ema = True
    ddpm.ema_start, ddpm.ema_update_rate = 0, 5          # the reference's rate (every 5th step), gate open from the start
    B, M, Cd = args.batch, NET["input_dim"], NET["cond_dim"]
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.rand(B, Cd, generator=g).to(dev)
    y = (torch.rand(B, M, generator=g) * (20.0 / 40.0)).to(dev)      # row sums ~ W = 20
    torch.manual_seed(1234 + rank)                                  # per-rank RNG stream for ts / noise / mask

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    warm = max(args.warmup, 10)
    steps = 50 if args.steps == 3 else args.steps                    # SURVEY cfg5: 50 timed steps after 10 warm-up
    # this library's kernels per step (tcgen05 forward / backward nodes, Adam+EMA), counted on one eager step with
    # lr = 0: inside the timed region they are replayed from CUDA graphs, which the launch counter cannot see
    lr0 = tr.opt.lr
    tr.opt.set_lr(0.0)
    _lib.launch_count(reset=True)
    tr._fwd_bwd(y, x)
    tr.opt.step()
    own_kernels_per_step = _lib.launch_count()
    tr.opt.reset_state()
    tr.opt.set_lr(lr0)
    for _ in range(warm):
        loss = tr.step(y, x)
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    with ClockSampler(local_rank) as clk:
        ev[0].record()
        for _ in range(steps):
            loss = tr.step(y, x)
        ev[1].record()
        barrier()
    ms = torch.tensor([ev[0].elapsed_time(ev[1])], device=dev, dtype=torch.float64)
    same = True
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        flat = tr.flat.flat.clone()
        dist.broadcast(flat, 0)
        same = bool(torch.equal(flat, tr.flat.flat))
    ms = float(ms)
    if rank == 0:
        f_train = 3 * 2 * (566400 + 100480)                          # fwd + dgrad + wgrad of the x-path and cond GEMMs
        value = B * world * steps / (ms * 1e-3)
        pk, pk_kind = peaks()
        peak_tf = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
        line = {"metric": TRAIN_METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": steps, "warmup": warm,
                "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16x3",
                "data": "synthetic",
                "config": {"workload": "BASELINE configs[4]: 80c MSR eps-MSE training step + EMA every 5th step, synthetic "
                                       "x = rand(B,80), y = rand(B,80) * W/40, data-parallel: one flat-gradient NCCL all-reduce per step",
                           "batch_per_gpu": B, "T": T, "optimizer": "fused Adam + EMA kernel (diffsg_adam_step), lr 1e-3",
                           "graph": "eager" if args.no_graph else "forward+backward and optimiser replayed as CUDA graphs",
                           "gemm": "this library's tcgen05 kernels (tlin_fwd_kernel / tlin_bwd_kernel: bf16 hi+lo operands, fp32 TMEM "
                                   "accumulators, LayerNorm->Swish fused into the operand prologue / dgrad epilogue); no cuBLAS in the step"},
                "clocks": clk.summary(), "gpu_launches": own_kernels_per_step * steps, "own_kernels_per_step": own_kernels_per_step,
                "final_loss": float(loss),
                "replicas_identical": same, "allreduce_bytes_per_step": tr.flat.numel * 4 if world > 1 else 0,
                "roofline": {"bound": "tensor", "achieved": value / world * f_train / 1e12, "peak": peak_tf, "unit": "TFLOP/s",
                             "frac": value / world * f_train / 1e12 / peak_tf, "traffic": None, "flop_per_sample": f_train,
                             "peak_source": f"{pk_kind} bf16 sustained",
                             "note": "GEMMs of width <= 256 on 128-row tiles: bound by operand staging (fp32 -> bf16 hi+lo through the "
                                     "LSU) and launch latency, far from the tensor roofline; three MMAs per product are not counted"}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ sampling mode
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="sample", choices=["sample", "train"])
    ap.add_argument("--rows", type=int, default=1 << 20, help="condition rows per GPU per step")
    ap.add_argument("--ref-rows", type=int, default=8192, help="rows per step of the CPU reference arm (BASELINE.md: 512 and 8192)")
    ap.add_argument("--cpu-rows", type=int, default=4096, help="rows of the in-run cpu_baseline sample")
    ap.add_argument("--eager-rows", type=int, default=65536, help="rows per call of the stock-eager GPU baseline")
    ap.add_argument("--batch", type=int, default=8192, help="--mode train: per-GPU batch (SURVEY cfg5: 512 / 8192 / 65536)")
    ap.add_argument("--no-graph", action="store_true", help="--mode train: eager instead of CUDA-graph replay")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip strong / precision / parity / eager legs (profiling runs)")
    ap.add_argument("--global-stats", action="store_true",
                    help="N > 1: treat the shards as ONE batch (whole-batch statistics, one 24-byte all-reduce per re-normalised step)")
    ap.add_argument("--precision", default="auto", choices=["auto", "fp32", "fp16x2", "fp16x3"])
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    if args.mode == "train":
        run_train(args, rank, local_rank, world)
        return

    import torch.distributed as dist
    from diffsg_b200 import _lib
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ddpm = build_model(dev)
    ddpm.model.precision = args.precision
    ddpm.noise_mode = "philox"          # in-kernel noise: no host traffic on the device-resident path
    B, M, Cd = args.rows, NET["input_dim"], NET["cond_dim"]
    ddpm.philox_seed = rank              # one Philox key per shard: streams stay disjoint however many calls each rank makes
    ddpm.philox_offset = 0               # (sample() advances the row offset by B per call)
    g = torch.Generator().manual_seed(rank)
    cond_host = torch.rand(B, Cd, generator=g).pin_memory()     # U(0,1) = min-max scaled gains
    cond = cond_host.to(dev, non_blocking=True)
    out_host = torch.empty(B, M).pin_memory()
    stats_group = dist.group.WORLD if (args.global_stats and world > 1) else None
    engine = ddpm.model.engine()
    x_macs, c_macs = engine.program.gemm_macs()
    engine_name = {"fp32": "simt-fp32 (CUDA cores)", "fp16x2": "tcgen05 fp16x2 (A hi+lo, W fp16, fp32 accum, one-MUFU Swish)",
                   "fp16x3": "tcgen05 fp16x3 (A hi+lo, W hi+lo, fp32 accum)"}[engine.precision]
    f_alg = T * 2 * 2 * x_macs + 2 * c_macs                     # 45.51 MFLOP for 80c

    def timed(fn, n):
        """n calls of fn between two events on the current stream, barrier + synchronize on both sides -> ms (max over ranks)."""
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(n):
            fn()
        ev[1].record()
        barrier()
        ms = torch.tensor([ev[0].elapsed_time(ev[1])], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    def step_resident():
        return ddpm.sample(cond, OMEGA, stats_group=stats_group)

    copy_stream = torch.cuda.Stream(device=dev)
    cond_bufs = [torch.empty(B, Cd, device=dev) for _ in range(2)]

    def run_e2e(n):
        """n steps through the public API with HOST buffers in and out, as a serving loop would run them: the
        pinned-host -> device copy of step k+1 and the device -> host read of step k's result are issued on a copy
        stream while step k / k+1 computes (double-buffered inputs).  Every byte still moves inside the timed region."""
        cur = torch.cuda.current_stream(dev)
        h2d = [torch.cuda.Event() for _ in range(n)]
        done = [None, None]

        def issue_h2d(k):
            with torch.cuda.stream(copy_stream):
                if done[k % 2] is not None:
                    copy_stream.wait_event(done[k % 2])           # the step that read this buffer has finished
                cond_bufs[k % 2].copy_(cond_host, non_blocking=True)
                h2d[k].record(copy_stream)

        copy_stream.wait_stream(cur)
        issue_h2d(0)
        for k in range(n):
            if k + 1 < n:
                issue_h2d(k + 1)
            cur.wait_event(h2d[k])
            y = ddpm.sample(cond_bufs[k % 2], OMEGA, stats_group=stats_group)
            ev_done = torch.cuda.Event()
            ev_done.record(cur)
            done[k % 2] = ev_done
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ev_done)
                out_host.copy_(y.reshape(B, M), non_blocking=True)
                y.record_stream(copy_stream)
        cur.wait_stream(copy_stream)

    for _ in range(args.warmup):
        step_resident()
    barrier()
    _lib.launch_count(reset=True)
    with ClockSampler(local_rank) as clk:
        ms = timed(step_resident, args.steps)
    launches = _lib.launch_count()
    # end to end through the public API, host buffers in and out
    run_e2e(1)
    ms2 = timed(lambda: run_e2e(args.steps), 1)

    # BASELINE's own configuration: ONE 1 Mi-row batch split over the N GPUs (strong scaling)
    strong = None
    if not args.no_extras:
        total_rows = 1 << 20
        if world == 1 and B == total_rows:
            strong = {"total_rows": total_rows, "rows_per_gpu": B, "value": B * args.steps / (ms * 1e-3), "ms_per_step": ms / args.steps,
                      "note": "identical to `value` at one GPU"}
        else:
            from diffsg_b200.parallel import shard_rows
            sl = shard_rows(total_rows, rank, world)
            gs = torch.Generator().manual_seed(4242)
            cond_s = torch.rand(total_rows, Cd, generator=gs)[sl].to(dev)
            ddpm.sample(cond_s, OMEGA, stats_group=stats_group)
            ms_s = timed(lambda: ddpm.sample(cond_s, OMEGA, stats_group=stats_group), args.steps)
            tiles, ctas = -(-cond_s.shape[0] // 128), max(engine.info().get("tc_grid_max", 296), 1)
            strong = {"total_rows": total_rows, "rows_per_gpu": cond_s.shape[0], "value": total_rows * args.steps / (ms_s * 1e-3),
                      "ms_per_step": ms_s / args.steps,
                      "note": f"{tiles} tiles of 128 rows on {ctas} persistent CTAs per GPU: {-(-tiles // ctas)} wave(s), "
                              f"the last one {100.0 * ((tiles - 1) % ctas + 1) / ctas:.0f} % full"}
            del cond_s

    if rank == 0:
        pk, pk_kind = peaks()
        total = B * world * args.steps
        value = total / (ms * 1e-3)
        e2e = total / (ms2 * 1e-3)
        peak_tf = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
        ach_tf = (B * args.steps * f_alg) / (ms * 1e-3) / 1e12     # per GPU (rank 0's kernels)
        traffic, traffic_note = None, "no ncu capture committed for this engine"
        if engine.precision != "fp32" and TRAFFIC_FILE.exists():
            tj = json.loads(TRAFFIC_FILE.read_text())
            traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) / (tj["rows"] * tj["steps"]) * B * T
            traffic_note = (f"ncu dram__bytes_read+write of the phase-B launch measured AT {tj['rows']} rows ({tj['steps']} reverse steps, "
                            f"{TRAFFIC_FILE.name}) per (row, step) x rows x T; algorithmic: {B * (Cd + M) * 4 + 4 * 2 * B * M * 4} B "
                            "(the rest is the per-CTA skip / cond scratch: 296 x ~0.5 MB exceeds the 126 MB L2)")
        line = {"metric": METRIC, "value": value, "unit": "solutions/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32" if engine.precision == "fp32" else "f16", "data": "synthetic",
                "config": sample_config(B, "whole batch (all-reduce of 2 doubles per re-normalised step)" if stats_group is not None
                                        else "per shard (reference per-call semantics)"),
                "engine": engine_name, "engine_info": engine.info(),
                "clocks": clk.summary(),
                "e2e": {"value": e2e, "unit": "solutions/s", "h2d_bytes_per_step": B * Cd * 4 * world,
                        "d2h_bytes_per_step": B * M * 4 * world, "ms_per_step": ms2 / args.steps},
                "gpu_launches": launches,
                "roofline": {"bound": "tensor", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s",
                             "frac": ach_tf / peak_tf, "traffic": traffic, "traffic_note": traffic_note,
                             "kernel": ("sample_simt_kernel" if engine.precision == "fp32" else "tc_unet_kernel<sampler>")
                                       + " (the 5 launches of one sample() call; > 99 % of the step's GPU time, profiles/r02_ncu_launches_tc.csv)",
                             "flop_per_solution": f_alg, "peak_source": f"{pk_kind} bf16 sustained"},
                "strong": strong}
        sd = {k: v.detach().cpu().clone() for k, v in ddpm.state_dict().items()}
        if world == 1 and not args.no_extras:
            if engine.precision == "fp16x2":      # the ~fp32 sibling mode on the same workload
                ddpm.model.precision = "fp16x3"
                ddpm.sample(cond, OMEGA)
                ms3 = timed(lambda: ddpm.sample(cond, OMEGA), 2)
                line["precision_modes"] = {"fp16x2": value, "fp16x3": B * 2 / (ms3 * 1e-3), "unit": "solutions/s",
                                           "note": "fp16x3 = activations AND weights as fp16 (hi, lo), exact two-MUFU Swish: ~fp32 accuracy"}
                ddpm.model.precision = args.precision
            line["parity"] = parity_record(ddpm, dev)
            eager = gpu_eager_run(sd, dev, args.eager_rows, 2)
            line["gpu_eager_baseline"] = (None if eager is None else
                                          {"value": eager, "unit": "solutions/s", "rows_per_call": args.eager_rows,
                                           "what": "UNMODIFIED reference DDPM.sample (baseline/_ref) on cuda:0, stock PyTorch eager, same weights / T / omega",
                                           "speedup_value": value / eager})
        if world == 1 and not args.no_cpu_baseline:
            torch.set_num_threads(host_threads())
            times, kind = cpu_reference_run(sd, args.cpu_rows, 2, 1)
            v = args.cpu_rows / (sum(times) / len(times))
            line["cpu_baseline"] = {"value": v, "unit": "solutions/s", "cores": torch.get_num_threads(), "kind": kind,
                                    "sample": f"{args.cpu_rows} rows x 2 repetitions after 1 warm-up, same net/T/omega, torch CPU"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
