#!/bin/bash
OUT=gpurun_out/r3; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_dp.py -m gpu -q -x 2>&1 | tail -5 > $OUT/pytest_gpu_dp_n2_final.log; cat $OUT/pytest_gpu_dp_n2_final.log
