python tools/prof_dd.py C4 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dd_ -c 3 -o gpurun_out/r1_c4_dd_v2 python tools/prof_dd.py C4 > gpurun_out/prof_ncu.log 2>&1
tail -2 gpurun_out/prof_ncu.log
