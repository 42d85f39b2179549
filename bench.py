#!/usr/bin/env python
"""Headline benchmark: CFG-DDPM solutions/s on the 80-channel MSR configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = one full pass of the hot path over one batch: `DDPM.sample` for `--rows`
synthetic condition rows (default 1 Mi per GPU; BASELINE.json configs[1]) = T=20 reverse
steps x 2 UNet passes + guidance + posterior update + 4 batch re-normalisations.
Rank 0 prints ONE JSON line (see the task contract): `value` = device-resident throughput,
`e2e` = the same through the public API with pinned-host inputs/outputs copied inside the
timed region, `roofline` = algorithmic FLOP/s of the sampler kernels against the measured
bf16 tensor peak, `cpu_baseline` = the reference algorithm (oracle port, torch CPU, all host
threads) on a bounded sample in the same run.

`--impl reference` times the reference's own CPU implementation of the path (the oracle port
in oracle/ddpm_oracle.py, bit-identical to the reference under make_golden.py) on the box's
host cores; under torchrun only rank 0 runs it.
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

T = 20
OMEGA = 500.0
NET = dict(input_dim=80, proj_dim=128, cond_dim=80, dims=(64, 32, 16, 8), is_attn=(False,) * 4,
           middle_attn=False, n_blocks=2)   # 80c = the 3c script with M=80 (ASSUMED)
METRIC = "CFG-DDPM solutions/sec (80c MSR, T=20, omega=500)"
FALLBACK_PEAKS = {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0}


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


def cpu_reference_run(sd, rows, reps, warmup):
    """Reference algorithm on the host cores: oracle port of DDPM.sample (torch CPU ops)."""
    from oracle import ddpm_oracle as O   # the one sanctioned use of oracle/ outside tests
    g = torch.Generator().manual_seed(0)
    cond = torch.rand(rows, NET["cond_dim"], generator=g)
    times = []
    with torch.no_grad():
        for r in range(warmup + reps):
            y_T, steps = O.draw_noise(rows, (1, NET["input_dim"]), T, 1000 + r)
            t0 = time.perf_counter()
            O.sample(sd, T, cond, OMEGA, y_T, steps)
            if r >= warmup:
                times.append(time.perf_counter() - t0)
    return times


def run_reference(args, rank):
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm is entitled to every host core
    try:
        torch.set_num_threads(len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        torch.set_num_threads(os.cpu_count() or 1)
    ddpm = build_model("cpu")
    sd = {k: v.detach().clone() for k, v in ddpm.state_dict().items()}
    rows = args.ref_rows
    times = cpu_reference_run(sd, rows, args.steps, args.warmup)
    dt = sum(times) / len(times)
    val = rows / dt
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "solutions/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"MSR 80c (assumed UNet1D 80/128/80/(64,32,16,8)/2) CFG sampling, {rows} rows/step on host CPU",
                       "T": T, "omega": OMEGA, "rows_per_step": rows},
            "cpu_baseline": {"value": val, "unit": "solutions/s", "cores": cores, "kind": "port",
                             "sample": f"{rows} rows x {args.steps} steps, torch CPU, reference algorithm (oracle port, bit-identical to reference DDPM.sample)"},
            "e2e": {"value": val, "unit": "solutions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=1 << 20, help="condition rows per GPU per step")
    ap.add_argument("--ref-rows", type=int, default=2048, help="rows per step of the CPU reference arm")
    ap.add_argument("--cpu-rows", type=int, default=4096, help="rows of the in-run cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    ddpm.philox_offset = rank * B        # disjoint noise streams per shard
    g = torch.Generator().manual_seed(rank)
    cond_host = torch.rand(B, Cd, generator=g).pin_memory()     # U(0,1) = min-max scaled gains
    cond = cond_host.to(dev, non_blocking=True)
    out_host = torch.empty(B, M).pin_memory()
    stats_group = dist.group.WORLD if (args.global_stats and world > 1) else None
    engine = ddpm.model.engine()
    x_macs, c_macs = engine.program.gemm_macs()
    engine_name = {"fp32": "simt-fp32 (CUDA cores)", "fp16x2": "tcgen05 fp16x2 (A hi+lo, W fp16, fp32 accum)",
                   "fp16x3": "tcgen05 fp16x3 (A hi+lo, W hi+lo, fp32 accum)"}[engine.precision]
    f_alg = T * 2 * 2 * x_macs + 2 * c_macs                     # 45.51 MFLOP for 80c

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
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    with ClockSampler(local_rank) as clk:
        ev[0].record()
        for _ in range(args.steps):
            step_resident()
        ev[1].record()
        barrier()
    launches = _lib.launch_count()
    ms = torch.tensor([ev[0].elapsed_time(ev[1])], device=dev, dtype=torch.float64)
    # end to end through the public API, host buffers in and out
    run_e2e(1)
    barrier()
    ev2 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev2[0].record()
    run_e2e(args.steps)
    ev2[1].record()
    barrier()
    ms2 = torch.tensor([ev2[0].elapsed_time(ev2[1])], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    ms, ms2 = float(ms), float(ms2)

    if rank == 0:
        pk, pk_kind = peaks()
        total = B * world * args.steps
        value = total / (ms * 1e-3)
        e2e = total / (ms2 * 1e-3)
        peak_tf = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
        ach_tf = (B * args.steps * f_alg) / (ms * 1e-3) / 1e12     # per GPU (rank 0's kernels)
        traffic = None                       # DRAM bytes of the kernels of one step, scaled from the committed ncu capture
        tpath = Path(__file__).resolve().parent / "profiles" / "r01_traffic_tc.json"
        if engine.precision != "fp32" and tpath.exists():
            tj = json.loads(tpath.read_text())
            traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) / (tj["rows"] * tj["steps"]) * B * T
        line = {"metric": METRIC, "value": value, "unit": "solutions/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32" if engine.precision == "fp32" else "f16", "data": "synthetic",
                "config": {"workload": "BASELINE configs[1]: MSR 80c CFG-DDPM sampling (assumed UNet1D 80/128/80/(64,32,16,8)/2, "
                                       "init_weights N(0,0.01) seed 0), synthetic rand(B,80) conditions, rows sharded per GPU, no collective",
                           "rows_per_gpu": B, "T": T, "omega": OMEGA, "noise": "in-kernel Philox4x32-10",
                           "batch_stats": "whole batch (all-reduce of 2 doubles per re-normalised step)" if stats_group is not None
                           else "per shard (reference per-call semantics)",
                           "l2": f"inputs+state per step {2 * B * M * 4 / 2**20:.0f} MiB > 126 MB L2; no explicit flush",
                           "engine": engine_name, "engine_info": engine.info()},
                "clocks": clk.summary(),
                "e2e": {"value": e2e, "unit": "solutions/s", "h2d_bytes_per_step": B * Cd * 4 * world,
                        "d2h_bytes_per_step": B * M * 4 * world, "ms_per_step": ms2 / args.steps},
                "gpu_launches": launches,
                "roofline": {"bound": "tensor", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s",
                             "frac": ach_tf / peak_tf, "traffic": traffic,
                             "traffic_note": "bytes per step per GPU = ncu dram read+write per (row, reverse step) of the phase-B launch "
                                             "(profiles/r01_traffic_tc.json) x rows x T; not measured in this run",
                             "kernel": ("sample_simt_kernel" if engine.precision == "fp32" else "tc_unet_kernel<true>")
                                       + " (the 5 launches of one sample() call; 99.8 % of the step's GPU time, profiles/r01_ncu_launches_tc.csv)",
                             "flop_per_solution": f_alg, "peak_source": f"{pk_kind} bf16 sustained"}}
        if world == 1 and not args.no_cpu_baseline:
            sd = {k: v.detach().cpu().clone() for k, v in ddpm.state_dict().items()}
            times = cpu_reference_run(sd, args.cpu_rows, 2, 1)
            v = args.cpu_rows / (sum(times) / len(times))
            line["cpu_baseline"] = {"value": v, "unit": "solutions/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"{args.cpu_rows} rows x 2 repetitions after 1 warm-up, same net/T/omega, torch CPU"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
