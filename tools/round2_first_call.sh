#!/usr/bin/env bash
# First GPU call of the next round: validate and time the experiments prepared (opt-in) at the end of round 1.
#   gpurun --timeout 600 -- 'bash tools/round2_first_call.sh > gpurun_out/round2_first.log 2>&1; cat gpurun_out/round2_first.log'
mkdir -p gpurun_out
timeout 60 tools/bin/ubench_pipes | tee gpurun_out/ubench_pipes_r2.txt          # tick calibration + VIMNMX / FFMA2 mixes
for v in 6 7 8 9 10 12; do
  echo "== FSR1_EASU_QUAD_VARIANT=$v  (6 = default; 7 = f32x2-packed per-pixel analysis; 8 = 7 + integer distance clamp; 9 = 8 + predicate-free interior tiles + incremental tile coordinates; 10 = 9 with scalar fp32 analysis; 12 = 9 at 7 CTAs per SM)"
  FSR1_EASU_QUAD_VARIANT=$v timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q \
      -k "fp16_kernels or golden or end_to_end or full_size or slabs or pipeline or flat" 2>&1 | tail -2
  FSR1_EASU_QUAD_VARIANT=$v timeout 90 python tools/variant_time.py 2x 2>&1 | tail -2
  FSR1_EASU_QUAD_VARIANT=$v timeout 90 python tools/pipeline_time.py 2>&1 | tail -1
done
for prio in "0,-1" "-1,0"; do
  echo -n "stream priorities easu,rcas = $prio: "
  FSR1_PIPE_PRIO=$prio timeout 90 python tools/pipeline_time.py 2>&1 | tail -1
done
for v in 0 1; do
  echo "== FSR1_EASU_PAIRS_VARIANT=$v (any-scale kernel; 1 = packed fp32 analysis + factored distance + integer clamp)"
  FSR1_EASU_PAIRS_VARIANT=$v timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "fp16_kernels or golden or end_to_end or slabs or dynamic" 2>&1 | tail -1
  for s in 1.5x 1.3x; do FSR1_EASU_PAIRS_VARIANT=$v timeout 90 python tools/variant_time.py $s 2>&1 | tail -1; done
done
for v in 0 1; do
  echo "== FSR1_UNORM_TILED=$v (RGBA8 images: 0 = direct kernels, 1 = TMA-tiled EASU + packed RCAS)"
  FSR1_UNORM_TILED=$v timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -k unorm 2>&1 | tail -1
  FSR1_UNORM_TILED=$v timeout 120 python tools/unorm_time.py 2>&1 | tail -2
done
for v in 0 1; do
  echo "== FSR1_RCAS_F32_VARIANT=$v (RGBA32F RCAS: 1 = MUFU reciprocal instead of IEEE __frcp_rn, -25 % instructions)"
  FSR1_RCAS_F32_VARIANT=$v timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "fp32_default or flat" 2>&1 | tail -1
  FSR1_RCAS_F32_VARIANT=$v timeout 120 python bench.py --workload 1080p-4k-fp32 --no-cpu --no-pipeline --steps 100 --warmup 10 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:(round(v['us'],1)) for k,v in d['kernels'].items()})"
done
