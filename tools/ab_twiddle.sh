#!/bin/sh
# A/B of twiddle generation variants (needs build/ab/*.so).
for v in default r1 r2; do
  if [ $v = default ]; then unset BBT_B200_LIB; else export BBT_B200_LIB=$PWD/build/ab/libbbt_$v.so; fi
  echo "== $v"
  python tests/accuracy.py
  python bench.py --steps 8 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('C2', d['value'], d['ms_per_step'], d.get('kernels'))"
  python bench.py --workload C4 --steps 8 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('C4', d['value'], d['ms_per_step'], d.get('kernels'))"
done
