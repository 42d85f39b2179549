#!/usr/bin/env python
"""Summarise an `ncu --set full --import-source on` report of tc_unet_kernel: headline raw metrics, warp-stall
reasons of the epilogue warps, and the hottest source lines.  Usage: tools/ncu_stalls.py gpurun_out/tc.ncu-rep > out.md"""
import collections
import csv
import io
import subprocess
import sys


def ncu_csv(rep, *args):
    out = subprocess.run(["ncu", "-i", rep, "--csv", *args], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep):
    raw = ncu_csv(rep, "--page", "raw")
    d = dict(zip(raw[0], raw[2] if len(raw) > 2 else raw[1]))
    units = dict(zip(raw[0], raw[1])) if len(raw) > 2 else {}
    keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
            "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
            "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
    print(f"# ncu summary of `{d.get('Kernel Name', 'kernel')}`\n")
    print("| metric | value | unit |\n|---|---|---|")
    for k in keys:
        if k in d:
            print(f"| `{k}` | {d[k]} | {units.get(k, '')} |")
    rows = ncu_csv(rep, "--page", "source", "--print-source", "cuda,sass")
    per = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    text, cur, hdr = {}, None, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
            si, ei = hdr.index("# Samples"), hdr.index("Instructions Executed")
            st = [(h, i) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        elif r[0].isdigit() and hdr:
            key = (cur, int(r[0]))
            text[key] = r[1].strip()
            try:
                per[key][0] += int(r[si]); per[key][1] += int(r[ei])
                for h, i in st:
                    per[key][2][h] += int(r[i] or 0)
            except ValueError:
                pass
    tot = sum(v[0] for v in per.values())
    agg = collections.Counter()
    for v in per.values():
        agg.update(v[2])
    print(f"\n## warp-stall samples by reason (all warps, {tot} samples)\n\n| reason | share |\n|---|---|")
    for k, n in agg.most_common(10):
        print(f"| {k} | {100 * n / max(tot, 1):.1f} % |")
    print("\n`stall_sleep` is the parked TMA / MMA lanes (`mbarrier.try_wait` with a suspend hint) and the idle lanes of the two "
          "producer warps at the final barrier; the rest is the four epilogue warps.\n")
    print("## hottest source lines\n\n| file:line | samples | warp instructions | top stalls | source |\n|---|---|---|---|---|")
    for (f, ln), (n, ex, stc) in sorted(per.items(), key=lambda kv: -kv[1][0])[:25]:
        top = ", ".join(f"{k[6:]} {100 * c / max(n, 1):.0f}%" for k, c in stc.most_common(2))
        print(f"| {f}:{ln} | {100 * n / max(tot, 1):.2f} % | {ex} | {top} | `{text[(f, ln)][:80].replace('|', '/')}` |")


if __name__ == "__main__":
    main(sys.argv[1])
