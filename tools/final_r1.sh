#!/usr/bin/env bash
# Round-end validation + evidence, most important first; every step under its own timeout.
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/final_gpu_tests.log; cat gpurun_out/final_gpu_tests.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/final_smoke.log
timeout 150 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; tail -c 600 gpurun_out/final_bench_n1.json; tail -2 gpurun_out/final_bench_n1.err
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/final_ncu_bench.log 2>&1; wc -l gpurun_out/final_launches.csv
timeout 150 ncu --set full --clock-control none --import-source on -k regex:'easu_h|rcas_h' -s 4 -c 2 -f -o gpurun_out/final_prof python tools/profile_run.py 2x 4 > gpurun_out/final_ncu_full.log 2>&1
timeout 60 ncu -i gpurun_out/final_prof.ncu-rep --page raw --csv > gpurun_out/final_ncu_full_raw.csv 2>/dev/null
timeout 60 python tools/ncu_summary.py gpurun_out/final_prof.ncu-rep > gpurun_out/final_ncu_summary.txt 2>&1; grep -c "=====" gpurun_out/final_ncu_summary.txt
for knobs in "0 4" "1 4" "0 8" "1 8"; do
  set -- $knobs
  echo "== pointwise layout=$1 n=$2"
  FSR1_POINT_LAYOUT=$1 FSR1_POINT_N=$2 timeout 60 python -m pytest tests/test_gpu_pointwise.py -m gpu -q 2>&1 | tail -1
  FSR1_POINT_LAYOUT=$1 FSR1_POINT_N=$2 timeout 60 python tools/pointwise_time.py --half-only gpurun_out/pointwise_l$1_n$2.json 2>&1 | grep -o "'op': '[a-z0-9>_-]*'\|'us': [0-9.]*\|'frac_of_hbm_peak': [0-9.]*" | paste - - - 
done
