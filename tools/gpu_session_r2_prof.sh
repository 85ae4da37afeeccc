#!/bin/bash
# round 2, 1-GPU profiling session: launch list of the bench command, one --set full capture of every bench kernel, A/B runs
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_at_size.py -q -x -k "allreduce" 2>&1 | tail -3
( time timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err ) 2>&1 | tail -3
( time timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_r02 \
    python tools/run_profile_r02.py > gpurun_out/prof_r02.log 2>&1 ) 2>&1 | tail -3
tail -3 gpurun_out/prof_r02.log
echo "== A/B permute TMA"; timeout 300 python tools/ab_permute_tma.py 2>&1 | tail -12
echo "== A/B pack lines (default)"; timeout 200 python tools/ab_c64.py 2>&1 | grep EINSUM
echo "== A/B pack 4-byte writers"; MB200_PACK=permute timeout 200 python tools/ab_c64.py 2>&1 | grep EINSUM
echo "== bench"; ( time timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err ) 2>&1 | tail -3
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r02_n1.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "lat", d.get("tiny_contraction_latency_us"))
for k, v in d["per_config"].items():
    print(k, round(v["value"], 2), v["clocks"])
for k, v in d["roofline_k1"].items():
    print(k, round(v["achieved"]), round(v["frac"], 3), v["clocks"])
PY
