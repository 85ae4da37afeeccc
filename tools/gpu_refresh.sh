#!/bin/bash
# Round-end measurement pass on one B200 (run under gpurun), in stages so that each call's gpurun_out/ stays below the
# 64 MiB copy-back limit:   bash tools/gpu_refresh.sh bench | ncu1 | ncu2 | ncu3
# bench: bench.py (both arms), per-kernel tables, probes, ncu launch list.  ncu1 / ncu2: `--set full` captures per kernel
# family. tools/ncu_summarise.py turns the outputs into profiles/.
set -x
mkdir -p gpurun_out
NCU="ncu --clock-control none"
case "${1:-bench}" in
bench)
  python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
  tail -c 600 gpurun_out/bench_n1.err
  python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
  python tools/bench_kernels.py > gpurun_out/kernels.log 2>&1
  python tools/bench_kernels.py --family > gpurun_out/kernels_family.log 2>&1
  python tools/bench_kernels.py --svd > gpurun_out/kernels_svd.log 2>&1
  python tools/bench_network.py > gpurun_out/network.log 2>&1
  python tools/bench_midsize.py > gpurun_out/kernels_midsize.log 2>&1
  python tools/probe_split_scheme.py > gpurun_out/split_scheme.log 2>&1
  for w in cfg3 c64 f32 c128; do python tools/power_probe.py $w; done > gpurun_out/power_probe.log 2>&1
  python tools/ab_c64.py > gpurun_out/ab_c64.log 2>&1
  python tools/ab_gather_pack.py > gpurun_out/ab_gather_pack.log 2>&1
  $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
  ;;
ncu1)
  $NCU --set full --import-source on -k regex:'gett_kernel|stream_kernel' -c 3 -f -o gpurun_out/prof_gett python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1
  $NCU --set full --import-source on -k regex:'tf32_gemm|permute' -c 9 -f -o gpurun_out/prof_tf32 python tools/run_cfg3_once.py > gpurun_out/ncu_tf32.log 2>&1
  ;;
ncu2)
  $NCU --set full --import-source on -k regex:permute -c 4 -f -o gpurun_out/prof_permute python tools/run_permute_once.py > gpurun_out/ncu_perm.log 2>&1
  $NCU --set full --import-source on -k regex:'hadamard|unary' -c 6 -f -o gpurun_out/prof_family python tools/run_family_once.py > gpurun_out/ncu_family.log 2>&1
  ;;
ncu3)
  $NCU --set full --import-source on -k regex:'persistent|pack_gather|splitk_reduce|gett_kernel|tf32_gemm' -c 10 -f -o gpurun_out/prof_extra python tools/run_extra_once.py > gpurun_out/ncu_extra.log 2>&1
  ;;
esac
ls -la gpurun_out | tail -30
du -sh gpurun_out
