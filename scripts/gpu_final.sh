#!/bin/bash
# Round-end evidence on ONE B200: GPU test suite, smoke, phase tables, bench (N = 1) and the reference arm.
OUT=gpurun_out/r3
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > $OUT/pytest_gpu_final.log; cat $OUT/pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_final.log 2>&1; tail -3 $OUT/smoke_final.log
for nb in 256 128 64 32; do timeout 120 python scripts/phase_times.py c4 $nb > $OUT/phase_c4_${nb}_final.txt 2>&1; tail -n 1 $OUT/phase_c4_${nb}_final.txt; done
for wl in c1 c2 c3; do timeout 120 python scripts/phase_times.py $wl > $OUT/phase_${wl}_final.txt 2>&1; tail -n 1 $OUT/phase_${wl}_final.txt; done
timeout 600 python bench.py > $OUT/bench_c4_n1_final.json 2> $OUT/bench_c4_n1_final.err; tail -c 600 $OUT/bench_c4_n1_final.json
