#!/bin/sh
# A/B of the C4 row-pass variants (third session; results in
# profiles/r2c_ab_row_variants.md).  Each variant's output is compared with the
# first one's by tools/ab.py.
python -m pytest tests/test_kernels.py -m gpu -x -q -k "row_variants or dedisperse_large" 2>&1 | tail -2
AB_FRAMES=32 python tools/ab.py C4 base row_landp=1 row_landp=2 row_landp=0 \
    dd_hint=267 dd_hint=8459 dd_hint=8459,row_landp=2 dd_hint=0,row_landp=0 \
    row2=0 row2=0,row_e16=1 row2=1,row_e16=0 2>&1 | tee gpurun_out/ab_row_variants.log
