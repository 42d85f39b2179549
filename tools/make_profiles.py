#!/usr/bin/env python
"""Turn the round-2 captures under gpurun_out/ (tools/profile_r2.sh, tools/sanitize_r2.sh, tools/scale_runs.sh) into
the tracked evidence under profiles/: raw ncu page, stall summary, measured DRAM traffic, launch list, scaling tables,
sanitizer logs, SASS instruction census of the in-tree library."""
import collections
import csv
import io
import json
import re
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
G, P = ROOT / "gpurun_out", ROOT / "profiles"


def ncu_csv(rep, *args):
    return subprocess.run(["ncu", "-i", str(rep), "--csv", *args], capture_output=True, text=True, check=True).stdout


def main():
    rep = G / "r2_tc_1mi.ncu-rep"
    if rep.exists():
        raw = ncu_csv(rep, "--page", "raw")
        (P / "r02_ncu_full_tc_unet_kernel_1mi.csv").write_text(raw)
        rows = list(csv.reader(io.StringIO(raw)))
        d = dict(zip(rows[0], rows[2]))
        u = dict(zip(rows[0], rows[1]))
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
        rd = float(d["dram__bytes_read.sum"]) * scale[u["dram__bytes_read.sum"]]
        wr = float(d["dram__bytes_write.sum"]) * scale[u["dram__bytes_write.sum"]]
        (P / "r02_traffic_tc.json").write_text(json.dumps({
            "kernel": d.get("Kernel Name"), "rows": 1048576, "steps": 16, "dram_bytes_read": rd, "dram_bytes_write": wr,
            "gpu_time_ms": float(d["gpu__time_duration.sum"]),
            "source": "ncu --set full --clock-control none, phase-B launch (reverse steps 15..0) of bench.py --rows 1048576 (tools/profile_r2.sh)"},
            indent=1) + "\n")
        md = subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_stalls.py"), str(rep)], capture_output=True, text=True).stdout
        (P / "r02_ncu_stalls_tc_unet_kernel_1mi.md").write_text(md)
    lc = G / "r2_launches.csv"
    if lc.exists():
        lines = [l for l in lc.read_text().splitlines() if l.startswith('"')]
        (P / "r02_ncu_launches_tc.csv").write_text("\n".join(lines) + "\n")
    for f in sorted(G.glob("r2_scale_*gpu.json")):
        txt = f.read_text().strip()
        if txt:
            shutil.copy(f, P / f.name.replace("r2_", "r02_"))
    # one table for the scaling runs
    table = []
    for n in (1, 2, 4, 8):
        row = {"n_gpus": n}
        f = G / f"r2_scale_sample_{n}gpu.json"
        if f.exists() and f.read_text().strip():
            j = json.loads(f.read_text().strip().splitlines()[-1])
            row.update(sample_weak=j["value"], sample_e2e=j["e2e"]["value"], sample_strong=(j.get("strong") or {}).get("value"))
        for b in (512, 8192, 65536):
            f = G / f"r2_scale_train_b{b}_{n}gpu.json"
            if f.exists() and f.read_text().strip():
                j = json.loads(f.read_text().strip().splitlines()[-1])
                row[f"train_b{b}"] = j["value"]
                row[f"train_b{b}_ms"] = j["ms_per_step"]
        table.append(row)
    if any(len(r) > 1 for r in table):
        (P / "r02_scaling_summary.json").write_text(json.dumps(table, indent=1) + "\n")
    for name in ("r2_memcheck_smoke.log", "r2_racecheck_smoke.log", "r2_synccheck_smoke.log", "r2_ubench.log", "r2_tanh_err.log"):
        if (G / name).exists():
            txt = (G / name).read_text()
            txt = "\n".join(l for l in txt.splitlines() if "Host Frame" not in l)[-20000:]
            (P / name.replace("r2_", "r02_")).write_text(txt + "\n")
    # ---- tcgen05 training kernels (tools/train_profile.sh)
    for kind in ("fwd", "bwd"):
        rep = G / f"t_{kind}.ncu-rep"
        if rep.exists():
            (P / f"r02_ncu_full_tlin_{kind}_kernel.csv").write_text(ncu_csv(rep, "--page", "raw"))
    lc = G / "t_launches.csv"
    if lc.exists():
        lines = [l for l in lc.read_text().splitlines() if l.startswith('"')]
        (P / "r02_ncu_launches_train.csv").write_text("\n".join(lines) + "\n")
    for tool in ("memcheck", "racecheck", "synccheck"):
        f = G / f"t_{tool}_tlin.log"
        if f.exists():
            txt = "\n".join(l for l in f.read_text().splitlines() if "Host Frame" not in l)[-20000:]
            (P / f"r02_{tool}_tlin.log").write_text(txt + "\n")
    for b in (512, 65536):
        f = G / f"timeline_{b}.txt"
        if f.exists():
            (P / f"r02_train_timeline_b{b}.txt").write_text("\n".join(l for l in f.read_text().splitlines() if "grid=" in l) + "\n")
    # SASS census of the library that ships
    so = ROOT / "diffsg_b200" / "libdiffsg_b200.so"
    if so.exists():
        sass = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True).stdout
        cnt, cur = collections.defaultdict(collections.Counter), None
        for line in sass.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                cur = m.group(1)
                continue
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
            if m and cur:
                cnt[cur][m.group(1)] += 1
        keys = ["UTCHMMA", "UTCBAR", "LDTM", "UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "MUFU", "F2FP", "HADD2", "STS", "LDS", "LDG", "STG", "RED", "ACQBULK"]
        out = ["# SASS census of diffsg_b200/libdiffsg_b200.so (cuobjdump -sass), tensor-core kernels",
               "", "| kernel | instructions | " + " | ".join(keys) + " |", "|---|---|" + "---|" * len(keys)]
        for fn, c in cnt.items():
            if "tc_unet_kernel" in fn or "tc_gemm_test" in fn or "tlin_" in fn:
                out.append(f"| `{fn}` | {sum(c.values())} | " + " | ".join(str(c[k]) for k in keys) + " |")
        out += ["", "UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk (1-D TMA), SYNCS = mbarrier ops,",
                "FFMA2 / FADD2 / FMUL2 = packed f32x2 arithmetic, MUFU = ex2 / rcp / tanh / rsqrt."]
        (P / "r02_sass_census.md").write_text("\n".join(out) + "\n")
    print("profiles updated:", sorted(p.name for p in P.glob("r02_*")))


if __name__ == "__main__":
    main()
