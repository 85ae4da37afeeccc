#!/bin/bash
# round 2, 8-GPU session: fused all-reduce policy at N = 8, then on-hardware parity and the bench line
N=${1:-8}
mkdir -p gpurun_out
export MB200_DIST_TIMELINE=1
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 tools/diag_allreduce.py 2>&1 | grep DIAG
  MB200_DIST_OVERLAP=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29622 tools/diag_allreduce.py 2>&1 | grep DIAG
  MB200_DIST_OVERLAP=0 MB200_DIST_MULTICAST=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29623 tools/diag_allreduce.py 2>&1 | grep DIAG
) | tee gpurun_out/diag_allreduce_n${N}.log | cut -c1-700
unset MB200_DIST_TIMELINE
bash tools/gpu_session_r2_n.sh $N 5 2>&1 | tail -25 | cut -c1-1500
