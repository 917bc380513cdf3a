ncu --query-metrics 2>/dev/null | grep -i "^nvl" | awk '{print $1}' > gpurun_out/nvl_metrics.txt; wc -l gpurun_out/nvl_metrics.txt; head -30 gpurun_out/nvl_metrics.txt
python tools/nvlink_epilogue.py 2>&1 | tail -2
M=$(grep -E "^nvl(rx|tx)__bytes(_data_user)?$" gpurun_out/nvl_metrics.txt | sed 's/$/.sum/' | paste -sd, -)
echo "metrics: $M"
if [ -n "$M" ]; then timeout 300 ncu --metrics $M,gpu__time_duration.sum --clock-control none -k regex:k_skyvis_finalize -s 3 -c 1 --csv --log-file gpurun_out/nvlink_epilogue_r02.csv python tools/nvlink_epilogue.py > gpurun_out/nvlink_epilogue.out 2>&1; tail -8 gpurun_out/nvlink_epilogue_r02.csv | cut -d, -f5,10-15; fi
timeout 300 python bench.py --config 4 --steps 5 --warmup 3 > gpurun_out/bench_r02_config4.json 2> gpurun_out/bench_r02_config4.err; tail -3 gpurun_out/bench_r02_config4.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r02_config4.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['tile_beam'], d['e2e']['precision_report'])"
