#!/bin/bash
# quick GPU loop: short bench (device-resident) + one ncu pass with DRAM bytes for k_decode
tag=$1
timeout 120 python bench.py --no-cpu --no-e2e --no-online --no-deskew > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$tag.json"))
print("BENCH Gpts/s %.1f ms/step %.3f k_decode_ms %.3f frac %.3f whole %.3f" % (d["value"]/1e9, d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["roofline"]["frac_whole_step"]))
PY
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_decode -s 3 -c 1 --csv --log-file gpurun_out/ncu_$tag.csv python bench.py --packets 262144 --steps 2 --warmup 3 --no-cpu --no-e2e --no-online --no-deskew > /dev/null 2>&1
grep -v "^==" gpurun_out/ncu_$tag.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin): print('NCU', r['Metric Name'], r['Metric Unit'], r['Metric Value'])
"
