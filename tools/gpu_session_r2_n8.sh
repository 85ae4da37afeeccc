#!/bin/bash
# round 2, 8-GPU session: fused all-reduce at N GPUs (policy variants), then on-hardware parity and the bench line
N=${1:-8}
mkdir -p gpurun_out
export MB200_DIST_TIMELINE=1
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 tools/diag_allreduce.py 2>&1 | grep DIAG
  true
  true
  MB200_DIST_MULTICAST=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29624 tools/diag_allreduce.py 2>&1 | grep DIAG
) | tee gpurun_out/diag_allreduce_n${N}.log | cut -c1-700
unset MB200_DIST_TIMELINE
if [ "${2:-full}" = "full" ]; then bash tools/gpu_session_r2_n.sh $N 5 2>&1 | tail -12 | cut -c1-600; fi
