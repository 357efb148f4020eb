# Round-2 (fourth session) evidence run on the B200 for the final build; each
# ncu command only after the same command line exited 0 without ncu.
# Outputs: gpurun_out/ (raw pages exported on the box; the reports stay there).
set -x
# 1. launch list of the driver's default bench (C4): shares of GPU time
python bench.py --steps 2 --warmup 3 --no-cpu --no-strong > gpurun_out/r2d_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/r2d_c4_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-strong > gpurun_out/r2d_ncu_launch.log 2>&1
# 2. C3 with the padded work buffer: one launch of each hot kernel
python bench.py --workload C3 --steps 1 --warmup 3 --no-cpu > gpurun_out/r2d_c3_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:'pfb_pair|dd_row|dd_col' \
    -s 12 -c 4 -o /tmp/r2d_c3_full python bench.py --workload C3 --steps 1 --warmup 3 --no-cpu > gpurun_out/r2d_ncu_c3.log 2>&1
ncu -i /tmp/r2d_c3_full.ncu-rep --page raw --csv > gpurun_out/r2d_c3_raw.csv
du -sh gpurun_out
