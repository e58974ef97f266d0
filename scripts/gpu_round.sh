#!/usr/bin/env bash
# development: one GPU call = world-edge + projected-kernel parity, the secondary workload lines, the fp32 mode, the headline line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_world_edges_gpu.py tests/test_gpu_parity.py -x -q -m gpu -k "world or golden or cloud or radius or degenerate or million or projected" > gpurun_out/r_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r_pytest.log
HGN_BENCH_NO_ROLLOUT=1 timeout 300 python bench.py --workload cfg2 --steps 5 > gpurun_out/r_cfg2.json 2> gpurun_out/r_cfg2.err; echo "cfg2 rc=$?"
HGN_BENCH_NO_ROLLOUT=1 timeout 300 python bench.py --workload cfg4 --steps 5 > gpurun_out/r_cfg4.json 2> gpurun_out/r_cfg4.err; echo "cfg4 rc=$?"
timeout 300 python scripts/fp32_mode.py > gpurun_out/r_fp32.json 2> gpurun_out/r_fp32.err; echo "fp32 rc=$?"
timeout 600 python bench.py --steps 3 > gpurun_out/r_cfg5.json 2> gpurun_out/r_cfg5.err; echo "cfg5 rc=$?"
python - <<'PY'
import json
for f in ("r_cfg2", "r_cfg4", "r_cfg5"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"] / 1e6, 1), "M/s", round(d["ms_per_step"], 2), "ms; e2e", round(d["e2e"]["value"] / 1e6, 1), "cpu", round(d["cpu_baseline"]["value"] / 1e6, 3),
              "torch-cuda", d.get("torch_cuda_reference"))
    except Exception as e:
        print(f, "unreadable", e)
print(open("gpurun_out/r_fp32.json").read()[:600])
PY
