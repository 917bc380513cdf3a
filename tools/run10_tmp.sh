timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_tests5.log; cat gpurun_out/r2_tests5.log
for v in l2on l2off; do
PB200_LIB=build/var/lib_$v.so timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:^k_skyvis$ -s 3 -c 1 --csv --log-file gpurun_out/skyvis_dram_$v.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1; tail -3 gpurun_out/skyvis_dram_$v.csv | cut -d, -f13-15
done
rm -f gpurun_out/r2_err_variants4.txt
for v in l2on l2off; do PB200_LIB=build/var/lib_$v.so timeout 200 python tools/err_c2.py sorted:0.97 2>&1 | grep -v Warn | head -1 >> gpurun_out/r2_err_variants4.txt; done
for spc in 1 4; do PB200_SKYVIS_SPC=$spc PB200_LIB=build/var/lib_l2on.so timeout 200 python tools/err_c2.py sorted:0.97 2>&1 | grep -v Warn | head -1 | sed "s/^/spc=$spc /" >> gpurun_out/r2_err_variants4.txt; done
cat gpurun_out/r2_err_variants4.txt
