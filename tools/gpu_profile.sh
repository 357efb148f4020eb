# ncu captures of bench.py (each only after the same command exited 0 without ncu)
python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_launch.log 2>&1
python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'dd_|chanpow' -s 12 -c 4 -o gpurun_out/r1_c2_full python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
