#!/usr/bin/env bash
# First GPU call of the next round (everything here was prepared without a GPU):
#   1. the gather-rate probe: TMA gather4 vs per-row ld.global, random and sorted indices, box {64,1} and {64,4}
#   1b. the M = 64 accumulator layout / rate probe (two 64-row tiles in flight: DESIGN.md s8)
#   2. the experimental TMA-gather backward kernel: bitwise parity against the default kernel, under a timeout (a hang must not cost the box)
#   3. its timing against the default kernel on the cfg5 layer
# usage: gpurun --timeout 900 -- scripts/round2_first_call.sh
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o scripts/probes/gather_rate scripts/probes/gather_rate.cu -lcuda > gpurun_out/r2_probe_build.log 2>&1
timeout 120 scripts/probes/gather_rate 1000000 5992002 0 > gpurun_out/r2_gather_random.log 2>&1; echo "probe random rc=$?"; tail -12 gpurun_out/r2_gather_random.log
timeout 120 scripts/probes/gather_rate 1000000 5992002 1 > gpurun_out/r2_gather_sorted.log 2>&1; echo "probe sorted rc=$?"; tail -8 gpurun_out/r2_gather_sorted.log
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I hyper-graph-nets_b200/csrc -o scripts/probes/m64_layout scripts/probes/m64_layout.cu >> gpurun_out/r2_probe_build.log 2>&1
timeout 60 scripts/probes/m64_layout > gpurun_out/r2_m64_layout.log 2>&1; echo "m64 layout rc=$?"; tail -14 gpurun_out/r2_m64_layout.log
for box in 1 4; do
  HGN_GATHER4_BOX_ROWS=$box HGN_TEST_EXPERIMENTAL=1 timeout 180 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k tma_gather > gpurun_out/r2_g4_pytest_box$box.log 2>&1
  echo "g4 parity (box rows $box) rc=$?"; tail -3 gpurun_out/r2_g4_pytest_box$box.log
done
timeout 120 python scripts/time_edge.py 2> /dev/null | tail -1
HGN_EDGE_BWD_TMA_GATHER=1 timeout 120 python scripts/time_edge.py 2> gpurun_out/r2_g4_time.err | tail -1
HGN_EDGE_BWD_TMA_GATHER=1 HGN_TC_ABLATE=64 timeout 120 python scripts/time_edge.py 2> gpurun_out/r2_g4_timeline.log | tail -1
HGN_BENCH_NO_TORCH_REFERENCE=1 timeout 400 python bench.py --workload cfg3 --steps 5 > gpurun_out/r2_cfg3.json 2> gpurun_out/r2_cfg3.err; echo "cfg3 rc=$?"; tail -c 600 gpurun_out/r2_cfg3.json; tail -3 gpurun_out/r2_cfg3.err
HGN_TEST_EXPERIMENTAL=1 timeout 180 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k receiver_sorted > gpurun_out/r2_sorted_pytest.log 2>&1; echo "sorted storage parity rc=$?"; tail -3 gpurun_out/r2_sorted_pytest.log
HGN_EDGE_STORAGE=receiver_sorted HGN_BENCH_NO_ROLLOUT=1 HGN_BENCH_NO_TORCH_REFERENCE=1 timeout 400 python bench.py --steps 3 > gpurun_out/r2_sorted_bench.json 2> gpurun_out/r2_sorted_bench.err; echo "sorted storage bench rc=$?"
python -c "import json;d=json.loads(open('gpurun_out/r2_sorted_bench.json').read().strip().splitlines()[-1]);print('receiver-sorted storage:',round(d['value']/1e6,1),'M/s',{k['name']:round(k['ms_per_step'],2) for k in d['kernels']})"
HGN_TEST_EXPERIMENTAL=1 timeout 180 python -m pytest tests/test_rollout_gpu.py -x -q -m gpu -k graphed > gpurun_out/r2_graphed_pytest.log 2>&1; echo "graphed training step rc=$?"; tail -3 gpurun_out/r2_graphed_pytest.log
