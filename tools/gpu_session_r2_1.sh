#!/bin/bash
# round 2, GPU session 1 (1 GPU): test-suite, bench line, launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/s1_gpu.txt 2>&1
free -g >> gpurun_out/s1_gpu.txt; nproc >> gpurun_out/s1_gpu.txt
( time timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 ) > gpurun_out/s1_pytest.log 2>&1
( time timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/s1_bench_n1.json 2> gpurun_out/s1_bench_n1.err ) >> gpurun_out/s1_pytest.log 2>&1
tail -c 1500 gpurun_out/s1_bench_n1.err
echo "---- bench head"; head -c 1500 gpurun_out/s1_bench_n1.json
