#!/bin/bash
# launch lists (ncu, serialised) of the bench workload's shards at the final commit + a memcheck pass over the tests
# that exercise the wide feature extractor and the programmatic launches
OUT=gpurun_out/r3; mkdir -p $OUT
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/launches_c4_nb256.csv python scripts/one_iter.py c4 256 4 > $OUT/ncu_nb256.log 2>&1
python profiles/summarize_launches.py $OUT/launches_c4_nb256.csv 4 > $OUT/launches_c4_nb256_summary.txt; head -8 $OUT/launches_c4_nb256_summary.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/launches_c4_nb32.csv python scripts/one_iter.py c4 32 4 > $OUT/ncu_nb32.log 2>&1
python profiles/summarize_launches.py $OUT/launches_c4_nb32.csv 4 > $OUT/launches_c4_nb32_summary.txt; head -8 $OUT/launches_c4_nb32_summary.txt
timeout 500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_configs.py -m gpu -q -k "c4_shard32 or c1_mnist_b128 or c2_resisc45" > $OUT/sanitizer_memcheck_final.log 2>&1; echo rc=$? >> $OUT/sanitizer_memcheck_final.log; tail -6 $OUT/sanitizer_memcheck_final.log
