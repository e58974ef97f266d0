#!/usr/bin/env bash
# development: A/B of the edge kernels (old build vs new), then the -m gpu suite file by file
TAG=${1:-r2d}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 500 python scripts/ab_edge.py > gpurun_out/${TAG}_ab_edge.log 2>&1; echo "ab rc=$?"; cat gpurun_out/${TAG}_ab_edge.log | cut -c1-400
for f in tests/test_gpu_parity.py tests/test_rollout_gpu.py tests/test_training_graph_gpu.py tests/test_world_edges_gpu.py tests/test_scale_gpu.py tests/test_connector_gpu.py tests/test_dropin_gpu.py; do
  b=$(basename $f .py)
  s=$(date +%s)
  timeout 400 python -m pytest $f -v -s -m gpu --durations=8 -p no:cacheprovider > gpurun_out/${TAG}_$b.log 2>&1
  echo "$b rc=$? $(( $(date +%s) - s )) s: $(grep -E '(passed|failed|skipped|error)' gpurun_out/${TAG}_$b.log | tail -1 | cut -c1-200)"
done
grep -hE "^(FAILED|ERROR)" gpurun_out/${TAG}_test_*.log | head -30 | cut -c1-300
