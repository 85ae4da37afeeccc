#!/bin/bash
N=${1:-2}
timeout 300 python -m pytest tests/test_at_size.py -q -x -k "allreduce" 2>&1 | tail -3
export MB200_DIST_TIMELINE=1
for R in default 8 24; do
  if [ "$R" != "default" ]; then export MB200_DIST_REDUCER_SMS=$R; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 tools/diag_allreduce.py 2>&1 | grep DIAG
done
unset MB200_DIST_REDUCER_SMS
MB200_DIST_OVERLAP=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29622 tools/diag_allreduce.py 2>&1 | grep DIAG
MB200_DIST_MULTICAST=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29623 tools/diag_allreduce.py 2>&1 | grep DIAG
