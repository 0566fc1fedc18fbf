#!/bin/bash
# Developer helper: config 5 at size, and the unmodified drop-in path (benchmark tool, 16 threads) against software.
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 python bench.py --workload random4g --level 1 --steps 5 --warmup 3 --no-sweep --no-ratio --no-cpu --e2e-steps 2 > gpurun_out/${TAG}_random4g.json 2> gpurun_out/${TAG}_random4g.err; echo "random4g rc=$?"
python - <<'PY'
import sys; sys.path.insert(0, 'tools')
import corpus
data, label, info = corpus.load()
sub = b"".join(data[o:o + (1 << 17)] for o in range(0, len(data), 6 * (1 << 17)))[:32 << 20]
open('/tmp/bench32.bin', 'wb').write(sub)
print(len(sub))
PY
B=tools/qzstd_benchmark
for args in "-m0 -t16" "-m1 -t16" "-m1 -t64" "-m1 -t16 -B"; do
  echo "== $args"; timeout 300 $B $args -l3 -c128K -L3 -E1 /tmp/bench32.bin 2>&1 | tail -2
done > gpurun_out/${TAG}_dropin.log 2>&1
for args in "-m1 -t16" "-m1 -t64"; do
  echo "== QZSTD_COALESCE=1 $args"; QZSTD_COALESCE=1 timeout 300 $B $args -l3 -c128K -L3 -E1 /tmp/bench32.bin 2>&1 | tail -2
done >> gpurun_out/${TAG}_dropin.log 2>&1
cat gpurun_out/${TAG}_dropin.log
