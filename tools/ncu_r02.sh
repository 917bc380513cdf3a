#!/bin/bash
# Round-2 ncu captures (run under gpurun, one GPU).  Full-set captures replay the kernel ~40 times with a device-memory
# save/restore per pass, so they use reduced source counts (same kernels, same tile shapes, > 148 output tiles);
# the launch list and the DRAM byte counts (one pass) are taken on the headline command.
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_r02.out 2>&1
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:^k_skyvis$ -s 3 -c 1 --csv --log-file gpurun_out/skyvis_dram_final.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_skyvis$ -s 3 -c 1 -f -o gpurun_out/skyvis_r02 python bench.py --nsrc 40000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_skyvis.out 2>&1
if [ "$1" == "all" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_skyvis_fp64$ -s 1 -c 1 -f -o gpurun_out/skyvis_fp64_taper_r02 python bench.py --config 3 --nside 128 --steps 1 --warmup 3 --no-e2e > gpurun_out/ncu_fp64.out 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_delay_fft -s 2 -c 1 -f -o gpurun_out/delay_fft_w32_r02 python tools/perf_dt.py > gpurun_out/ncu_dt.out 2>&1
fi
ls -la gpurun_out/*.ncu-rep
