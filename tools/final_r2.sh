#!/usr/bin/env bash
# Round-end validation on one B200: tests, smoke, bench lines, config table, launch list, sanitizer.  Everything lands in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tools/final_r2.sh > gpurun_out/final_r2.log 2>&1; tail -30 gpurun_out/final_r2.log'
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_gpu_tests.log 2>&1; tail -3 gpurun_out/r02_gpu_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee -a gpurun_out/r02_gpu_tests.log
timeout 300 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/bench.err; tail -c 400 gpurun_out/bench.err
timeout 200 python bench.py --no-cpu --fused > gpurun_out/r02_bench_n1_fused.json 2>> gpurun_out/bench.err
timeout 200 python bench.py --no-cpu --workload 1080p-4k-unorm8 > gpurun_out/r02_bench_n1_unorm8.json 2>> gpurun_out/bench.err
timeout 200 python bench.py --no-cpu --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_k20.json 2>> gpurun_out/bench.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_n1*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), "Mpix/s", round(d["ms_per_step"] * 1e3, 2), "us", {k: round(v["us"], 1) for k, v in d["kernels"].items() if "us" in v},
              "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 3), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
timeout 500 python tools/run_configs.py > gpurun_out/r02_configs.json 2> gpurun_out/configs.err; tail -c 300 gpurun_out/configs.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_n1.csv python bench.py --steps 3 --warmup 3 --no-cpu > /dev/null 2>&1
timeout 400 compute-sanitizer --tool memcheck python tools/sanitize_run.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1; tail -3 gpurun_out/r02_sanitizer_memcheck.log
timeout 500 compute-sanitizer --tool racecheck python tools/sanitize_run.py > gpurun_out/r02_sanitizer_racecheck.log 2>&1; tail -3 gpurun_out/r02_sanitizer_racecheck.log
