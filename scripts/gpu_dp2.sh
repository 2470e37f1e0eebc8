#!/bin/bash
OUT=gpurun_out/r3; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q -x 2>&1 | tail -5 > $OUT/pytest_gpu_dp_n2_final.log; cat $OUT/pytest_gpu_dp_n2_final.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/scale_c4_n2_final.json 2> $OUT/scale_c4_n2_final.err
tail -c 400 $OUT/scale_c4_n2_final.json; tail -3 $OUT/scale_c4_n2_final.err
