rm -f gpurun_out/r2_err_variants5.txt
for v in h0 h1; do for spc in 2 4; do PB200_SKYVIS_SPC=$spc PB200_LIB=build/var/lib_$v.so timeout 200 python tools/err_c2.py sorted:0.97 2>&1 | grep -v Warn | head -1 | sed "s/^/spc=$spc /" >> gpurun_out/r2_err_variants5.txt; done; done
cat gpurun_out/r2_err_variants5.txt
for spc in 2 4; do
PB200_SKYVIS_SPC=$spc PB200_LIB=build/var/lib_h1.so timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:^k_skyvis$ -s 3 -c 1 --csv --log-file gpurun_out/skyvis_dram_h1_spc$spc.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1; grep -E "dram|time" gpurun_out/skyvis_dram_h1_spc$spc.csv | awk -F'","' '{print $(NF-2), $NF}'
done
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; tail -3 gpurun_out/r2_bench3.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench3.json')); print(d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['precision_report'], d['gather_check'])"
timeout 300 python bench.py --config 5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c5b.json 2> gpurun_out/r2_bench_c5b.err; tail -3 gpurun_out/r2_bench_c5b.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_c5b.json')); print(d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['tail'])"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
