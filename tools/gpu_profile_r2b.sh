# Round-2 (second session) evidence run on the B200; each ncu command only
# after the same command line exited 0 without ncu.  Outputs: gpurun_out/.
# The reports embed the whole module (> 64 MiB with four kernels, more than one
# gpurun call brings back), so their raw pages are exported as CSV on the box
# and the reports themselves are left there.
set -x
# 1. launch list of the driver's default bench (C4): shares of GPU time
python bench.py --steps 2 --warmup 3 --no-cpu --no-strong > gpurun_out/r2b_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/r2b_c4_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-strong > gpurun_out/r2b_ncu_launch.log 2>&1
# 2. full capture of one launch of each hot kernel of C4 (8 frames per launch)
AB_FRAMES=8 python tools/ab.py C4 base > gpurun_out/r2b_ab_plain.log 2>&1 && \
AB_FRAMES=8 ncu --set full --clock-control none -k regex:'dd_|chanpow' \
    -s 8 -c 4 -o /tmp/r2b_c4_full python tools/ab.py C4 base > gpurun_out/r2b_ncu_c4.log 2>&1
ncu -i /tmp/r2b_c4_full.ncu-rep --page raw --csv > gpurun_out/r2b_c4_raw.csv
# 3. the same for C3 (filter bank, interleaved dedispersion with Power fused in)
python bench.py --workload C3 --steps 1 --warmup 3 --no-cpu > gpurun_out/r2b_c3_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:'pfb|dd_' \
    -s 8 -c 4 -o /tmp/r2b_c3_full python bench.py --workload C3 --steps 1 --warmup 3 --no-cpu > gpurun_out/r2b_ncu_c3.log 2>&1
ncu -i /tmp/r2b_c3_full.ncu-rep --page raw --csv > gpurun_out/r2b_c3_raw.csv
du -sh gpurun_out
