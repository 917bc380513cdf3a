timeout 300 python -m pytest tests -m gpu -x -q -k "pipeline or drain or config5 or noise or snapshot" 2>&1 | tail -3
rm -f gpurun_out/r2_err_variants6.txt
for rep in 1 2; do for v in h0 h1b; do PB200_LIB=build/var/lib_$v.so timeout 200 python tools/err_c2.py sorted:0.97 2>&1 | grep -v Warn | head -1 >> gpurun_out/r2_err_variants6.txt; done; done
cat gpurun_out/r2_err_variants6.txt
PB200_LIB=build/var/lib_h1b.so timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:^k_skyvis$ -s 3 -c 1 --csv --log-file gpurun_out/skyvis_dram_h1b.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1; grep -E "dram|time" gpurun_out/skyvis_dram_h1b.csv | awk -F'","' '{print $(NF-2), $NF}'
timeout 300 python bench.py --config 5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c5c.json 2> gpurun_out/r2_bench_c5c.err; tail -3 gpurun_out/r2_bench_c5c.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_c5c.json')); print(d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['tail'])"
