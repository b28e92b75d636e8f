set -x
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3
timeout 250 python bench.py 2>&1 | tail -1 > gpurun_out/bench_final.json
python -c "
import json;d=json.load(open('gpurun_out/bench_final.json'));print(d['value'],d['ms_per_step'],d['e2e'],d['roofline']['frac'],d['roofline']['kernel_ms_per_step'],d['stage_ms_per_step'],d['roofline_hbm']['frac'],d['cpu_baseline']['value'],d['clocks'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file gpurun_out/launches_r1_final2.csv python bench.py --steps 2 --warmup 3 --no-graph > gpurun_out/b_ncu.log 2>&1
tail -2 gpurun_out/b_ncu.log | cut -c1-300
