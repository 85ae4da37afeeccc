"""Turns gpurun_out/prof_r02.ncu-rep + launches_r02.csv into the tracked round-2 summaries (run HERE, no GPU):
    python tools/ncu_summarise_r02.py
Writes profiles/ncu_r02_summary.md, profiles/launches_r02.csv, and updates profiles/ncu_traffic.json.
The kernel order inside prof_r02 is fixed by tools/run_profile_r02.py."""
import csv
import io
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
SC = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg.per_second",
    "launch__registers_per_thread", "launch__waves_per_multiprocessor", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]
# what each profiled launch is (order of tools/run_profile_r02.py) and its algorithmic bytes
LABELS = [
    ("cfg4b gett (the bench headline kernel, 16384^3)", 12884901888.0, "gett_z_cfg4b_dram_bytes_per_launch"),
    ("cfg4b split tail (60 tiles x 2 k-slices)", None, None),
    ("cfg4b split tail: ordered slice reduce", None, None),
    ("cfg1 gett (rank-4 dim 64, scrambled)", 805306368.0, "gett_z_cfg1_dram_bytes_per_launch"),
    ("cfg2 step 2a gett", 436207616.0, "gett_z_128x64_dram_bytes_per_launch"),
    ("cfg2 step 2b stream_kernel", 536875008.0, "stream_2b_dram_bytes_per_launch"),
    ("cfg2 step 2c gett", 436207616.0, None),
    ("cfg3 pack A (line writer)", 805306368.0, None), ("cfg3 pack B (line writer)", 805306368.0, None),
    ("cfg3 tcgen05 GEMM (CTA pairs)", None, None),
    ("cfg5 pack A", 402653184.0, None), ("cfg5 pack B", 402653184.0, None), ("cfg5 tcgen05 GEMM", None, None),
    ("K1 c128 64^4 kilj->ijkl, register-tile kernel", 536870912.0, None), ("K1 same, TMA-staged kernel", 536870912.0, None),
    ("K1 c64 (256,8,8,256,8)->(3,1,0,2,4), register-tile kernel", 536870912.0, None), ("K1 same, TMA-staged kernel", 536870912.0, None),
]


def main():
    os.makedirs(PROF, exist_ok=True)
    rep = os.path.join(OUT, "prof_r02.ncu-rep")
    md = ["# ncu summaries - round 2\n",
          "`ncu --set full --clock-control none --import-source on --profile-from-start off python tools/run_profile_r02.py` under `gpurun` on one "
          "B200 (`tools/gpu_session_r2_prof.sh`); read here without a GPU by `tools/ncu_summarise_r02.py`. Durations under ncu are "
          "cold-cache and serialised: compare shares and ratios, not absolutes; bench numbers come from `bench.py` without a profiler.\n"]
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(PROF, "ncu_traffic.json")))
    except Exception:
        pass
    if os.path.exists(rep):
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units, rows = rows[0], rows[1], rows[2:]
        idx = {h: i for i, h in enumerate(hdr)}
        md.append("| # | what | kernel | grid | ms | DRAM read + write (MB) | algorithmic MB | traffic / algorithmic | GB/s | tensor pipe % | issue % |\n|---|---|---|---|---|---|---|---|---|---|---|")
        cfg3 = cfg5 = 0.0
        for n, r in enumerate(rows):
            what, alg, key = LABELS[n] if n < len(LABELS) else ("?", None, None)
            t = float(r[idx["gpu__time_duration.sum"]])
            tu = units[idx["gpu__time_duration.sum"]]
            ms = t / 1e6 if tu.startswith("n") else (t / 1e3 if tu.startswith("u") else t)
            try:
                rd = float(r[idx["dram__bytes_read.sum"]]) * SC[units[idx["dram__bytes_read.sum"]]]
                wr = float(r[idx["dram__bytes_write.sum"]]) * SC[units[idx["dram__bytes_write.sum"]]]
            except Exception:
                rd = wr = float("nan")
            tot = rd + wr
            if key:
                traffic[key] = tot
            if 7 <= n <= 9:
                cfg3 += tot
            if 10 <= n <= 12:
                cfg5 += tot
            tp = max(float(r[idx[m]] or 0) for m in ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
                                                     "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"))
            name = r[idx["Kernel Name"]].replace("void ", "").replace("<unnamed>::", "")[:44]
            md.append(f"| {n} | {what} | `{name}` | {r[idx['Grid Size']]} | {ms:.4f} | {rd / 1e6:.1f} + {wr / 1e6:.1f} | "
                      f"{(alg / 1e6 if alg else float('nan')):.1f} | {(tot / alg if alg else float('nan')):.2f} | {tot / ms / 1e6:.0f} | {tp:.1f} | "
                      f"{float(r[idx['sm__issue_active.avg.pct_of_peak_sustained_elapsed']] or 0):.1f} |")
        traffic["cfg3_dram_bytes_per_call"] = cfg3
        traffic["cfg5_dram_bytes_per_call"] = cfg5
        for n, r in enumerate(rows):
            what = LABELS[n][0] if n < len(LABELS) else "?"
            md.append(f"\n### {n}: {what} - `{r[idx['Kernel Name']][:90]}`  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}\n")
            md.append("| metric | value | unit |\n|---|---|---|")
            for m in METRICS:
                if m in idx and r[idx[m]] not in ("", "n/a"):
                    md.append(f"| {m} | {r[idx[m]]} | {units[idx[m]]} |")
    lpath = os.path.join(OUT, "launches_r02.csv")
    if os.path.exists(lpath):
        rows = [r for r in csv.reader(open(lpath)) if len(r) > 10 and r[0].isdigit()]
        with open(os.path.join(PROF, "launches_r02.csv"), "w") as f:
            f.write("id,kernel,block,grid,gpu__time_duration_ns\n")
            for r in rows:
                f.write(f"{r[0]},\"{r[4][:160]}\",\"{r[7]}\",\"{r[8]}\",{r[-1]}\n")
        ours = [r for r in rows if "mb200" in r[4] or "unnamed>::" in r[4] and "at::" not in r[4]]
        tot = {}
        for r in rows:
            lib = "torch (synthetic data / parity helpers)" if ("at::" in r[4] or "elementwise" in r[4] or "index" in r[4].lower() and "at" in r[4]) else None
            key = lib or r[4].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[-70:]
            tot.setdefault(key, [0, 0.0])
            tot[key][0] += 1
            tot[key][1] += float(r[-1])
        allns = sum(v[1] for v in tot.values())
        md.append("\n## launch list of `bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e` (`launches_r02.csv`, first 400 launches)\n")
        md.append("| kernel | launches | total ms | share |\n|---|---|---|---|")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            md.append(f"| `{k}` | {v[0]} | {v[1] / 1e6:.3f} | {100 * v[1] / allns:.1f} % |")
    open(os.path.join(PROF, "ncu_r02_summary.md"), "w").write("\n".join(md) + "\n")
    json.dump(traffic, open(os.path.join(PROF, "ncu_traffic.json"), "w"), indent=1)
    print("\n".join(md[:30]))
    print(traffic)


if __name__ == "__main__":
    main()
