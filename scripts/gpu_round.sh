#!/usr/bin/env bash
# development: one single-GPU call = the whole -m gpu suite, the default bench line, cfg3, and the two ncu passes of the profiling
# recipe (launch list of one bench step; --set full of the edge kernels at the cfg5 size).  usage: gpurun -- bash scripts/gpu_round.sh [tag]
TAG=${1:-r2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; grep "^smoke" gpurun_out/${TAG}_smoke.log
timeout 900 python -m pytest tests -x -q -m gpu -s -p no:cacheprovider > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/${TAG}_pytest_gpu.log | cut -c1-250
grep -E "^(15 layers|slab|cfg5 full|dropin\[|PARTITION|.*rerouted)" gpurun_out/${TAG}_pytest_gpu.log | cut -c1-420
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/${TAG}_bench_n1.err | cut -c1-300
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_n1.json").read().strip().splitlines()[-1])
print("cfg5", round(d["value"] / 1e6, 1), "M/s", round(d["ms_per_step"], 2), "ms; e2e", round(d["e2e"]["value"] / 1e6, 1), "launches", d["gpu_launches"], "roofline", round(d["roofline"]["frac"], 3), "layer", round(d["layer_roofline"]["frac"], 3))
print({k["name"]: round(k["ms_per_step"], 2) for k in d["kernels"]}, "sum", round(sum(k["ms_per_step"] for k in d["kernels"]), 1))
print("rollout", {k: (round(v, 1) if isinstance(v, float) else (round(v["value"], 1) if isinstance(v, dict) and "value" in v else None)) for k, v in d.get("rollout", {}).items() if k in ("value", "cuda_graph", "cpu_reference")}, "cpu", d["cpu_baseline"]["kind"], round(d["cpu_baseline"]["value"]))
PY
HGN_BENCH_NO_TORCH_REFERENCE=1 timeout 400 python bench.py --workload cfg3 --steps 5 > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err; echo "cfg3 rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_cfg3.json").read().strip().splitlines()[-1])
print("cfg3", round(d["value"] / 1e6, 1), "M/s", round(d["ms_per_step"], 2), "ms", "graph", round(d["cuda_graph"]["value"] / 1e6, 1), "launches", d["gpu_launches"], {k["name"]: round(k["ms_per_step"], 2) for k in d["kernels"]})
PY
# ncu pass 1: launch list of one timed bench step (cold-cache, serialised: shares only)
HGN_BENCH_NO_ROLLOUT=1 HGN_BENCH_NO_TORCH_REFERENCE=1 HGN_BENCH_NO_CPU=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 1 --warmup 3 > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launch list rc=$?"
# ncu pass 2: --set full of the edge kernels at the cfg5 size (one launch each)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"edge_fwd_tc_kernel|edge_bwd_tc_kernel|segment_sum_bf16_128_kernel" -s 12 -c 6 -o gpurun_out/${TAG}_ncu_edge -f python scripts/time_edge.py > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"; ls -la gpurun_out/${TAG}_ncu_edge.ncu-rep 2>/dev/null
