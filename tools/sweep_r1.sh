#!/usr/bin/env bash
# One GPU call: new-code parity first, then the variant sweep (every step under its own timeout).
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_pointwise.py -m gpu -q 2>&1 | tail -4
timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "fp16_kernels or pipeline or end_to_end" 2>&1 | tail -2
for v in 2 6; do FSR1_EASU_QUAD_VARIANT=$v timeout 90 python tools/variant_time.py 2x 2>&1 | tail -2; done
for combo in "2 0 0" "6 0 0" "2 3 0" "6 3 0" "5 3 0" "5 0 0" "6 3 5"; do
  set -- $combo
  echo -n "easu=$1 rcas=$2 cap=$3: "
  FSR1_EASU_QUAD_VARIANT=$1 FSR1_RCAS_VARIANT=$2 FSR1_EASU_CTAS_PER_SM=$3 timeout 90 python tools/pipeline_time.py 2>&1 | tail -1
done
timeout 120 python tools/pointwise_time.py gpurun_out/pointwise_time.json 2>&1 | tail -14
