"""Turns gpurun_out/*.ncu-rep + launches.csv into the tracked summaries under profiles/ (run HERE, no GPU):
    python tools/ncu_summarise.py r01
Writes profiles/ncu_<round>_summary.md, profiles/launches_<round>.csv, profiles/ncu_traffic.json."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg.per_second",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_allocated", "launch__waves_per_multiprocessor",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
    "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def main():
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(PROF, exist_ok=True)
    md = [f"# ncu summaries — round {rnd}\n",
          "Captured under `gpurun` on one B200 with `ncu --set full --clock-control none --import-source on` "
          "(`tools/ncu_summarise.py` reads the `.ncu-rep` files here, without a GPU). Durations under ncu are "
          "cold-cache and serialised: compare shares, not absolutes; bench numbers come from `bench.py` without a profiler.\n"]
    traffic = {}
    for rep, title in (("prof_gett.ncu-rep", "gett kernels inside one bench step (2a, 2b, 2c)"),
                       ("prof_permute.ncu-rep", "K1 permute kernels (tools/bench_kernels.py, first cases)"),
                       ("prof_tf32.ncu-rep", "K3 tcgen05 3xTF32 kernels and their K1 split-writer packs (configs 3, 5; 8192^3 Float32)"),
                       ("prof_family.ncu-rep", "hadamard / unary_einsum streaming kernels"),
                       ("prof_extra.ncu-rep", "persistent gather-GEMM (8192 x 8192 x 256), split tail + ordered reduce (2048^3), tcgen05 on the last shape of tools/run_extra_once.py")):
        path = os.path.join(OUT, rep)
        if not os.path.exists(path):
            continue
        hdr, units, rows = raw(path)
        idx = {h: i for i, h in enumerate(hdr)}
        md.append(f"\n## {title} (`{rep}`)\n")
        for r in rows:
            name = r[idx["Kernel Name"]]
            md.append(f"\n### `{name}`  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}\n")
            md.append("| metric | value | unit |\n|---|---|---|")
            for m in METRICS:
                if m in idx and r[idx[m]] not in ("", "n/a"):
                    md.append(f"| {m} | {r[idx[m]]} | {units[idx[m]]} |")
            try:
                scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
                rd = float(r[idx["dram__bytes_read.sum"]]) * scale[units[idx["dram__bytes_read.sum"]]]
                wr = float(r[idx["dram__bytes_write.sum"]]) * scale[units[idx["dram__bytes_write.sum"]]]
                md.append(f"| **dram traffic (read+write)** | {(rd + wr) / 1e6:.1f} | MB |")
                if "CoreZ<128, 64" in name and "gett_z_128x64_dram_bytes_per_launch" not in traffic:
                    traffic["gett_z_128x64_dram_bytes_per_launch"] = rd + wr
                    traffic["gett_z_128x64_grid"] = r[idx["Grid Size"]]
                if "CoreZ<128, 16" in name:
                    traffic["gett_z_128x16_dram_bytes_per_launch"] = rd + wr
            except Exception:
                pass
    # launch list
    lpath = os.path.join(OUT, "launches.csv")
    if os.path.exists(lpath):
        rows = [r for r in csv.reader(open(lpath)) if len(r) > 10 and r[0].isdigit()]
        with open(os.path.join(PROF, f"launches_{rnd}.csv"), "w") as f:
            f.write("id,kernel,block,grid,gpu__time_duration_ns\n")
            for r in rows:
                f.write(f"{r[0]},\"{r[4]}\",\"{r[7]}\",\"{r[8]}\",{r[-1]}\n")
        tot = {}
        for r in rows:
            key = r[4].split("(")[0][-70:]
            tot.setdefault(key, [0, 0.0])
            tot[key][0] += 1
            tot[key][1] += float(r[-1])
        allns = sum(v[1] for v in tot.values())
        md.append(f"\n## launch list of `bench.py --steps 3 --warmup 3` (`launches_{rnd}.csv`)\n")
        md.append("| kernel | launches | total ms | share |\n|---|---|---|---|")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            md.append(f"| `{k}` | {v[0]} | {v[1] / 1e6:.3f} | {100 * v[1] / allns:.1f} % |")
    open(os.path.join(PROF, f"ncu_{rnd}_summary.md"), "w").write("\n".join(md) + "\n")
    if traffic:
        json.dump(traffic, open(os.path.join(PROF, "ncu_traffic.json"), "w"), indent=1)
    print("\n".join(md[-12:]))
    print(traffic)


if __name__ == "__main__":
    main()
