#!/usr/bin/env bash
# split / alternate / fused frame scheduling at N=1 and N=2, 200 and 20 steps (gpurun --gpus 2)
r1() { tag=$1; shift; timeout 100 python bench.py --no-cpu "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N1 $tag', round(d['value']), round(d['ms_per_step']*1e3,1), d.get('unpipelined_ms_per_step'))"; }
r2() { tag=$1; shift; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 "$@" 2>gpurun_out/err_$tag.log | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N2 $tag', round(d['value']), round(d['ms_per_step']*1e3,1), round(d['halo']['exposed_us_per_step'],1), round(d['halo']['us_per_step_without_exchange'],1), d['parity']['sharded_equals_single_gpu'], round(d['e2e']['value']))"; }
timeout 300 python -m pytest tests/test_gpu_sharding.py -m gpu -q 2>&1 | tail -2
r1 split200 --steps 200 --warmup 20; r1 alt200 --steps 200 --warmup 20 --alternate; r1 fused200 --steps 200 --warmup 20 --fused
r1 split20 --steps 20 --warmup 5; r1 alt20 --steps 20 --warmup 5 --alternate; r1 fused20 --steps 20 --warmup 5 --fused
r2 split200 --steps 200 --warmup 20; r2 alt200 --steps 200 --warmup 20 --alternate; r2 fused200 --steps 200 --warmup 20 --fused
r2 split20 --steps 20 --warmup 5; r2 alt20 --steps 20 --warmup 5 --alternate; r2 fused20 --steps 20 --warmup 5 --fused
