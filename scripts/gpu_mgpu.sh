#!/usr/bin/env bash
# usage: gpurun --gpus N -- bash scripts/gpu_mgpu.sh N
N=${1:-2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi topo -m 2>/dev/null | head -12
timeout 600 python -m pytest tests/test_partition_gpu.py -m gpu -q -s -x > gpurun_out/r2_partition_pytest_n$N.log 2>&1; echo "partition pytest rc=$?"; grep "PARTITION-GPU-OK\|passed\|failed\|skipped" gpurun_out/r2_partition_pytest_n$N.log | cut -c1-300; grep -B2 -A12 "Error\|error" gpurun_out/r2_partition_pytest_n$N.log | tail -30 | cut -c1-300
for mode in ${2:-peer nccl}; do
  HGN_HALO=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n${N}_$mode.json 2> gpurun_out/r2_bench_n${N}_$mode.err; echo "bench $mode rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n${N}_$mode.json').read().strip().splitlines()[-1])
    print('$mode', 'value %.1f M' % (d['value']/1e6), 'ms %.2f' % d['ms_per_step'], 'e2e %.1f M' % (d['e2e']['value']/1e6), 'kernel ms %.2f' % sum(k['ms_per_step'] for k in d['kernels']), 'check', d.get('partition_check'))
    print({k['name']: round(k['ms_per_step'],2) for k in d['kernels']})
except Exception as e:
    print('$mode: no line', e); print(open('gpurun_out/r2_bench_n${N}_$mode.err').read()[-1500:])
PY
done
