#!/usr/bin/env bash
# usage: gpurun --gpus 8 -- bash scripts/gpu_mgpu8.sh ["4:peer 8:peer 8:nccl"]     (N = 4 and N = 8 on one 8-GPU box)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_partition_gpu.py -m gpu -q -s > gpurun_out/r2_partition_pytest_n8.log 2>&1; echo "partition pytest rc=$?"; grep "PARTITION-GPU-OK rank 0\|passed\|failed" gpurun_out/r2_partition_pytest_n8.log | cut -c1-300
run() {  # N mode
  HGN_HALO=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $1 --steps 15 --warmup 3 > gpurun_out/r2_bench_n$1_$2.json 2> gpurun_out/r2_bench_n$1_$2.err; echo "bench N=$1 $2 rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n$1_$2.json').read().strip().splitlines()[-1])
    print('N=$1 $2', 'value %.1f M' % (d['value']/1e6), 'ms %.2f' % d['ms_per_step'], 'e2e %.1f M' % (d['e2e']['value']/1e6), 'kernel ms %.2f' % sum(k['ms_per_step'] for k in d['kernels']), 'check %.2e' % d['partition_check']['value'])
    print({k['name']: round(k['ms_per_step'],2) for k in d['kernels']})
except Exception as e:
    print('N=$1 $2: no line', e); print(open('gpurun_out/r2_bench_n$1_$2.err').read()[-1500:])
PY
}
for spec in ${1:-4:peer 8:peer 8:nccl}; do run ${spec%%:*} ${spec##*:}; done
