python tools/prof_dd.py C2 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dd_row -c 1 -o gpurun_out/r1_row_v2 python tools/prof_dd.py C2 > gpurun_out/prof_ncu.log 2>&1
tail -2 gpurun_out/prof_ncu.log
