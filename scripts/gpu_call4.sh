#!/usr/bin/env bash
# development: A/B of the edge-path kernels (library of the previous commit vs this tree), the -m gpu suite as the driver runs it, one bench line
TAG=${1:-r2f}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 400 python scripts/ab_edge.py > gpurun_out/${TAG}_ab_edge.log 2>&1; echo "ab rc=$?"; cat gpurun_out/${TAG}_ab_edge.log | cut -c1-400
s=$(date +%s)
timeout 900 python -m pytest tests -x -q -m gpu -s -p no:cacheprovider > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$? $(( $(date +%s) - s )) s"; tail -4 gpurun_out/${TAG}_pytest_gpu.log | cut -c1-250
grep -E "^bf16 input-gradient" gpurun_out/${TAG}_pytest_gpu.log | cut -c1-200
HGN_BENCH_NO_TORCH_REFERENCE=1 timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/${TAG}_bench_n1.err | cut -c1-300
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_n1.json").read().strip().splitlines()[-1])
print("cfg5", round(d["value"] / 1e6, 1), "M/s", round(d["ms_per_step"], 2), "ms; e2e", round(d["e2e"]["value"] / 1e6, 1), "launches", d["gpu_launches"], "roofline", round(d["roofline"]["frac"], 3), "layer", round(d["layer_roofline"]["frac"], 3))
print({k["name"]: round(k["ms_per_step"], 2) for k in d["kernels"]}, "sum", round(sum(k["ms_per_step"] for k in d["kernels"]), 1))
PY
