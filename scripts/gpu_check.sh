#!/bin/bash
# One GPU-box call: GPU test suite, phase tables at the latency-bound shapes, optional extras.
# usage: scripts/gpu_check.sh TAG [full] [notest]
TAG=${1:-x}
OUT=gpurun_out/r3
mkdir -p $OUT
if [ "$3" != "notest" ]; then
  timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $OUT/pytest_$TAG.log
  cat $OUT/pytest_$TAG.log
fi
timeout 120 python scripts/phase_times.py c4 32 > $OUT/phase_c4_32_$TAG.txt 2>&1
timeout 120 python scripts/phase_times.py c2 > $OUT/phase_c2_$TAG.txt 2>&1
if [ "$2" = "full" ]; then
  timeout 120 python scripts/phase_times.py c4 256 > $OUT/phase_c4_256_$TAG.txt 2>&1
  timeout 120 python scripts/phase_times.py c4 64 > $OUT/phase_c4_64_$TAG.txt 2>&1
fi
for f in $OUT/phase_*_$TAG.txt; do echo $f; tail -n 1 $f; done
