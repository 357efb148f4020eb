#!/bin/sh
# A/B of the C4 split 2048 x 8192 with the row pass in 256-thread CTAs.
python -m pytest tests/test_kernels.py -m gpu -x -q -k "dedisperse_large" 2>&1 | tail -2 > gpurun_out/t4_tests.log
cat gpurun_out/t4_tests.log
AB_FRAMES=32 python tools/ab.py C4 base dd_hint=267 dd_hint=8459 dd_hint=8459,row_landp=2 dd_hint=8459,row_landp=1 dd_hint=0,row_landp=0 dd_hint=8459 > gpurun_out/t4_ab.log 2>&1
cat gpurun_out/t4_ab.log
