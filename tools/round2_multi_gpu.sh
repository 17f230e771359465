#!/usr/bin/env bash
# Multi-GPU experiments prepared at the end of round 1 (run with gpurun --gpus N; N = 2 first, 8 to confirm):
#   gpurun --gpus 2 --timeout 600 -- 'bash tools/round2_multi_gpu.sh 2 > gpurun_out/round2_mgpu.log 2>&1; cat gpurun_out/round2_mgpu.log'
N=${1:-2}
mkdir -p gpurun_out
port=29520
for flags in "" "--halo-depth 3" "--halo-batch 4" "--pipeline-sharded" "--halo-batch 4 --pipeline-sharded"; do
  port=$((port + 1))
  echo "== N=$N bench.py $flags"
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus "$N" --steps 200 --warmup 20 --no-cpu $flags 2> gpurun_out/mgpu_err_$port.log | tail -1 |
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['parallelism'])"
done
