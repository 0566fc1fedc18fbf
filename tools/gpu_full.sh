#!/bin/bash
# Developer helper for one gpurun call: full GPU test suite, both bench arms, the ncu launch list and two full captures.
# Usage: tools/gpu_full.sh TAG
TAG=${1:-x}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 1 --no-sweep --no-ratio --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz77 -s 2 -c 1 -f -o gpurun_out/prof_${TAG}_L3 python tools/ncu_one.py 3 0 1617 > gpurun_out/${TAG}_ncu_L3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz77 -s 2 -c 1 -f -o gpurun_out/prof_${TAG}_L6 python tools/ncu_one.py 6 0 296 > gpurun_out/${TAG}_ncu_L6.log 2>&1
python tools/corpus_parts.py > gpurun_out/${TAG}_corpus_parts.txt 2>&1
ls -la gpurun_out | grep ${TAG}
