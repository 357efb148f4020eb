#!/bin/sh
python -m pytest tests/test_kernels.py -m gpu -x -q -k "row_variants" 2>&1 | tail -2 > gpurun_out/t7_tests.log
cat gpurun_out/t7_tests.log
AB_FRAMES=32 python tools/ab.py C4 base row2=0 row2=0,row_e16=1 row2=1,row_e16=0 row2=0,row_e16=1 > gpurun_out/t7_ab.log 2>&1
cat gpurun_out/t7_ab.log
