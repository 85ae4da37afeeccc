#!/bin/bash
timeout 300 python -m pytest tests/test_at_size.py -q -x -k "allreduce" 2>&1 | tail -3
export MB200_DIST_TIMELINE=1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 tools/diag_allreduce.py 2>&1 | grep DIAG | cut -c1-600
unset MB200_DIST_TIMELINE
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/_multi_gpu_worker.py 2>&1 | grep MULTI_GPU_REPORT | cut -c1-200
