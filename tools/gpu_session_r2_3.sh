#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_at_size.py tests/test_gpu.py -q -x -k "allreduce or tcgen05 or cfg3 or cfg5 or dangling or scatter" 2>&1 | tail -15 ) > gpurun_out/s3_pytest.log 2>&1
tail -15 gpurun_out/s3_pytest.log
echo "== A/B pack lines v2 (default)"; timeout 200 python tools/ab_c64.py 2>&1 | grep EINSUM
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__issue_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:pack_lines --csv python tools/run_cfg3_once.py 2>/dev/null | grep pack_lines | awk -F'","' '{print $5, $(NF-2), $(NF-1), $NF}' | head -24
