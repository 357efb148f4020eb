#!/bin/sh
python -m pytest tests/test_kernels.py tests/test_configs.py -m gpu -x -q -k "dedisperse or c3 or c4 or c5 or power_fused or detect" 2>&1 | tail -2 > gpurun_out/t6_tests.log
cat gpurun_out/t6_tests.log
AB_FRAMES=32 python tools/ab.py C4 base base > gpurun_out/t6_ab.log 2>&1
cat gpurun_out/t6_ab.log
python bench.py --workload C3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); print('C3', d['value'], d['ms_per_step'], {k:round(v['ms_per_launch'],3) for k,v in d['kernels'].items()})"
