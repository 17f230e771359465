#!/usr/bin/env bash
# Multi-GPU validation + measurement in one call (run with gpurun --gpus N):
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/multi_gpu_run.sh 8 > gpurun_out/mgpu_n8.log 2>&1; tail -40 gpurun_out/mgpu_n8.log'
# 1. the sharding tests that need several devices (IPC over NVLink, NCCL data plane, 4 ranks)
# 2. bench.py at N ranks: default (direct NVLink stores, two-kernel frames), the driver's 20-step run with device timestamps,
#    --halo nccl, --fused, and BASELINE configs[4] (2160p->8K cut into N slabs); every line carries "parity" (sharded == single GPU,
#    both data planes, oracle bands) and "halo".
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sharding.py -m gpu -q -k "${2:-two_processes or two_gpus or four_gpus}" 2>&1 | tail -2
port=29610
run() {
  port=$((port + 1))
  tag=$1; shift
  FSR1_TRACE_TAG=n${N}_$tag timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus "$N" "$@" > gpurun_out/bench_n${N}_${tag}.json 2> gpurun_out/bench_n${N}_${tag}.err
  python - "$tag" gpurun_out/bench_n${N}_${tag}.json <<'PY'
import json, sys
tag, path = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    print("%-12s %9.0f Mpix/s  %7.2f us/step  halo %s  parity %s  e2e %.0f" % (
        tag, d["value"], d["ms_per_step"] * 1e3,
        {k: (round(v, 2) if isinstance(v, float) else v) for k, v in (d.get("halo") or {}).items() if k in ("mode", "exposed_us_per_step", "us_per_step_without_exchange")},
        {k: v for k, v in (d.get("parity") or {}).items() if k != "oracle_bands"}, d["e2e"]["value"]))
except Exception as e:
    print(tag, "FAILED:", e)
    print(open(path.replace(".json", ".err")).read()[-1500:])
PY
}
run k200 --steps 200 --warmup 20
run k20 --steps 20 --warmup 5 --trace
run nccl --steps 200 --warmup 20 --halo nccl
run fused --steps 200 --warmup 20 --fused
run fused20 --steps 20 --warmup 5 --fused
run cfg5 --steps 200 --warmup 20 --workload 2160p-8k-fp16 --shard-frame
