#!/bin/bash
# Round-end measurement pass on one B200 (run under gpurun): bench (both arms), per-kernel table, ncu launch list
# and one `--set full` capture per kernel family. Outputs land in gpurun_out/; tools/ncu_summarise.py turns them
# into profiles/.
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 600 gpurun_out/bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python tools/bench_kernels.py > gpurun_out/kernels.log 2>&1
python tools/bench_kernels.py --family > gpurun_out/kernels_family.log 2>&1
python tools/bench_kernels.py --svd > gpurun_out/kernels_svd.log 2>&1
python tools/bench_network.py > gpurun_out/network.log 2>&1
python tools/bench_midsize.py > gpurun_out/kernels_midsize.log 2>&1
MB200_SPLITK=0 python tools/bench_midsize.py > gpurun_out/kernels_midsize_nosplit.log 2>&1
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
$NCU --set full --import-source on -k regex:'gett_kernel|stream_kernel' -c 3 -f -o gpurun_out/prof_gett python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_full.log 2>&1
$NCU --set full --import-source on -k regex:permute -c 4 -f -o gpurun_out/prof_permute python tools/run_permute_once.py > gpurun_out/ncu_perm.log 2>&1
$NCU --set full --import-source on -k regex:'tf32_gemm|permute' -c 9 -f -o gpurun_out/prof_tf32 python tools/run_cfg3_once.py > gpurun_out/ncu_tf32.log 2>&1
$NCU --set full --import-source on -k regex:'hadamard|unary' -c 6 -f -o gpurun_out/prof_family python tools/run_family_once.py > gpurun_out/ncu_family.log 2>&1
ls -la gpurun_out | tail -30
