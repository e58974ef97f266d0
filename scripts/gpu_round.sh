#!/usr/bin/env bash
# development: one GPU call = the whole -m gpu suite, the edge-kernel timing, a secondary workload line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r_pytest.log
timeout 120 python scripts/time_edge.py 2> gpurun_out/r_time_edge.err | tail -1
HGN_BENCH_NO_TORCH_REFERENCE=1 timeout 300 python bench.py --workload cfg2 --steps 5 > gpurun_out/r_cfg2b.json 2> gpurun_out/r_cfg2b.err; echo "cfg2 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r_cfg2b.json").read().strip().splitlines()[-1])
print("cfg2", d.get("cuda_graph"), round(d["value"] / 1e6, 1), "M/s", round(d["ms_per_step"], 2), "ms", {k["name"]: round(k["ms_per_step"], 3) for k in d["kernels"]})
PY
