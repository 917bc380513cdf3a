timeout 200 python -m pytest tests/test_gpu_api.py -m gpu -x -q -k "subband" 2>&1 | grep -E "Error|error|assert|^E " | head -20 > gpurun_out/r2_subband.log; cat gpurun_out/r2_subband.log
rm -f gpurun_out/r2_err_variants3.txt
for v in base gsync gsync_s1 gsync_st2 u8; do PB200_LIB=build/var/lib_$v.so timeout 200 python tools/err_c2.py sorted:0.97 2>&1 | grep -v Warn >> gpurun_out/r2_err_variants3.txt; done; cat gpurun_out/r2_err_variants3.txt
python tools/perf_dt.py > gpurun_out/r2_dt2.txt 2>&1; cat gpurun_out/r2_dt2.txt
timeout 300 python bench.py --config 3 --steps 3 --warmup 3 --no-e2e > gpurun_out/r2_bench_c3b.json 2> gpurun_out/r2_bench_c3b.err; tail -3 gpurun_out/r2_bench_c3b.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_c3b.json')); print(d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['healpix_gather'])"
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:^k_skyvis$ -s 3 -c 1 --csv --log-file gpurun_out/skyvis_dram_r02.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1; tail -4 gpurun_out/skyvis_dram_r02.csv
