#!/bin/bash
# round 2, final 1-GPU session: full GPU test-suite, the bench line, launch list + one --set full capture of every bench kernel, SVD timing
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) > gpurun_out/final_pytest.log 2>&1
tail -6 gpurun_out/final_pytest.log
( time timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err ) 2>&1 | tail -3
( time timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_reference_arm.json 2> /dev/null ) 2>&1 | tail -3
( time timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err ) 2>&1 | tail -3
( time timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_r02 \
    python tools/run_profile_r02.py > gpurun_out/prof_r02.log 2>&1 ) 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python tools/bench_kernels.py --svd 2>&1 | grep SVD | tee gpurun_out/kernels_svd_r02.log
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r02_n1.json"))
print("value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], "lat", d.get("tiny_contraction_latency_us"))
for k, v in d["per_config"].items():
    print(k, round(v["value"], 2), v["clocks"], v["parity"])
for k, v in d["roofline_k1"].items():
    print(k, round(v["achieved"]), round(v["frac"], 3))
print(open("gpurun_out/bench_r02_reference_arm.json").read()[:600])
PY
