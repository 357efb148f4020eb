# Round-2 evidence run on the B200 (each ncu command only after the same
# command line exited 0 without ncu).  Outputs land in gpurun_out/.
set -x
# 1. launch list of the driver's default bench (C4): shares of GPU time
python bench.py --steps 2 --warmup 3 --no-cpu --no-strong > gpurun_out/r2_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/r2_c4_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-strong > gpurun_out/r2_ncu_launch.log 2>&1
# 2. full capture of one launch of each hot kernel (8 frames per launch)
AB_FRAMES=8 python tools/ab.py C4 base > gpurun_out/r2_ab_plain.log 2>&1 && \
AB_FRAMES=8 ncu --set full --clock-control none --import-source on -k regex:'dd_|chanpow' \
    -s 8 -c 4 -o gpurun_out/r2_c4_full python tools/ab.py C4 base > gpurun_out/r2_ncu_full.log 2>&1
tail -2 gpurun_out/r2_ncu_full.log
