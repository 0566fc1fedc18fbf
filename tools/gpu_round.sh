#!/bin/bash
# Developer helper for one gpurun call: parity tests, then the timing loop.  Usage: tools/gpu_round.sh TAG [pytest-args]
TAG=${1:-x}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q "$@" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
timeout 600 python tools/quick_time.py all > gpurun_out/${TAG}_quick.log 2>&1; echo "quick rc=$?"
cat gpurun_out/${TAG}_quick.log
