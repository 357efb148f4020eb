#!/bin/sh
# Source-level (SASS) profile of one launch of the C4 row pass: where the
# warps' stall samples fall.  Only after the same command exited 0 without ncu.
AB_FRAMES=8 python tools/ab.py C4 base > gpurun_out/t3_plain.log 2>&1 && \
AB_FRAMES=8 ncu --set full --clock-control none --import-source on -k regex:dd_row2 \
    -s 3 -c 1 -o /tmp/t3_row python tools/ab.py C4 base > gpurun_out/t3_ncu.log 2>&1
ncu -i /tmp/t3_row.ncu-rep --page source --csv > gpurun_out/t3_row_src.csv 2> gpurun_out/t3_src.err
ncu -i /tmp/t3_row.ncu-rep --page raw --csv > gpurun_out/t3_row_raw.csv
ls -la /tmp/t3_row.ncu-rep gpurun_out/t3_row_src.csv
tail -3 gpurun_out/t3_ncu.log
