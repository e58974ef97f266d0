#!/usr/bin/env bash
# development: parity of the projected kernels, then the edge backward timing under a list of HGN_TC_ABLATE switch values
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "projected or interleaved" > gpurun_out/exp_pytest.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/exp_pytest.log
for ab in "$@"; do
  HGN_TC_ABLATE=$ab timeout 120 python scripts/time_edge.py 2> gpurun_out/exp_tl_$ab.log | tail -1
done
