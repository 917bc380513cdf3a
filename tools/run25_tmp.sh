rm -f gpurun_out/r2_err_variants10.txt
for rep in 1 2; do
for cfg in "auto 2" "recurrence_quarter 2" "recurrence_quarter 4" "auto 4"; do set -- $cfg; PB200_METHOD=$1 PB200_SKYVIS_SPC=$2 timeout 200 python tools/err_c2.py sorted:0.97 2>&1 | grep -v Warn | head -1 | sed "s/^/method=$1 spc=$2 /" >> gpurun_out/r2_err_variants10.txt; done
done
cat gpurun_out/r2_err_variants10.txt
