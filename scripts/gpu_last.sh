#!/bin/bash
OUT=gpurun_out/r3; mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > $OUT/pytest_gpu_last.log; cat $OUT/pytest_gpu_last.log
timeout 200 python bench.py > $OUT/bench_c4_n1_last.json 2> $OUT/bench_c4_n1_last.err; python - <<'PY'
import json
l=json.loads(open('gpurun_out/r3/bench_c4_n1_last.json').read().strip().splitlines()[-1])
r=l['roofline']; print('value', l['value'], 'ms', l['ms_per_step'], 'e2e', l['e2e']['value'], 'eval', l['eval']['value'], 'lstm us', r['launch_us'], 'frac', r['frac'], 'c2', l['secondary']['c2']['value'])
PY
