#!/bin/sh
# Source-level (SASS) profile of one launch of each C4 column pass and of the
# channelizer.  Only after the same command exited 0 without ncu.
AB_FRAMES=8 python tools/ab.py C4 base > gpurun_out/t5_plain.log 2>&1 && \
AB_FRAMES=8 ncu --set full --clock-control none --import-source on -k regex:'dd_col|chanpow' \
    -s 6 -c 3 -o /tmp/t5_col python tools/ab.py C4 base > gpurun_out/t5_ncu.log 2>&1
ncu -i /tmp/t5_col.ncu-rep --page source --csv > gpurun_out/t5_col_src.csv 2> gpurun_out/t5_src.err
ncu -i /tmp/t5_col.ncu-rep --page raw --csv > gpurun_out/t5_col_raw.csv
ls -la /tmp/t5_col.ncu-rep gpurun_out/t5_col_src.csv
tail -3 gpurun_out/t5_ncu.log
