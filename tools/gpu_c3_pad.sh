python -m pytest tests/test_kernels.py tests/test_configs.py tests/test_pfb_api.py -m gpu -x -q -k "dedisperse_large or c3 or pfb" 2>&1 | tail -2
for v in 1 0 1; do
BBT_TUNE=dd_work_pad=$v python bench.py --workload C3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); print('C3 pad=$v', round(d['value'],1), round(d['ms_per_step'],3), {k:round(v['ms_per_launch'],3) for k,v in d['kernels'].items()})"
done
