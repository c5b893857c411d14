#!/bin/bash
# final captures of the round: bench lines of the three workloads, launch list, full capture of the roofline kernel
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_composite.json 2> gpurun_out/bench_composite.err
timeout 600 python bench.py --workload surf --steps 5 --warmup 3 > gpurun_out/bench_surf.json 2> gpurun_out/bench_surf.err
timeout 900 python bench.py --workload stream8k --steps 5 --warmup 3 > gpurun_out/bench_stream8k.json 2> gpurun_out/bench_stream8k.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
for f in composite surf stream8k reference; do python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_$f.json') if l.startswith('{')][-1])
print('$f', 'value', round(d['value'],1), 'e2e', d['e2e']['value'], 'roof', (d.get('roofline') or {}).get('frac'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))"; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2_composite.csv python bench.py --steps 2 --warmup 3 --profile-mode > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:harris_fused3 -s 3 -c 1 -o gpurun_out/prof_r2_fused3_final python tools/harris_timing.py 48 0 > gpurun_out/ncu_fused3_final.log 2>&1
timeout 200 python tools/harris_timing.py 48 0 2>&1 | grep "^{" > gpurun_out/timing_final.txt; timeout 200 python tools/harris_timing.py 64 0 2>&1 | grep "^{" >> gpurun_out/timing_final.txt; timeout 200 python tools/harris_timing.py 64 1 2>&1 | grep "^{" >> gpurun_out/timing_final.txt
cut -c1-150 gpurun_out/timing_final.txt
timeout 300 python tools/front_timing.py > gpurun_out/front_timing.txt 2>&1
