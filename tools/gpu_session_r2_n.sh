#!/bin/bash
# round 2, multi-GPU session: on-hardware parity of the sharded / sliced paths, then the bench line at N GPUs
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/sN${N}_topo.txt 2>&1
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    tests/_multi_gpu_worker.py ) > gpurun_out/sN${N}_parity.log 2>&1
grep MULTI_GPU_REPORT gpurun_out/sN${N}_parity.log | sed 's/^MULTI_GPU_REPORT //' > gpurun_out/multi_gpu_parity_n${N}.json
tail -c 1500 gpurun_out/sN${N}_parity.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 \
    bench.py --gpus $N --steps ${2:-5} --warmup 3 > gpurun_out/bench_r02_n${N}.json 2> gpurun_out/bench_r02_n${N}.err ) 2>&1 | tail -3
tail -c 1200 gpurun_out/bench_r02_n${N}.err
echo "---- bench"; head -c 3000 gpurun_out/bench_r02_n${N}.json
