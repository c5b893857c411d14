#!/bin/bash
mkdir -p gpurun_out
for cfg in "64 0" "48 0"; do timeout 200 python tools/harris_timing.py $cfg; done > gpurun_out/timing.txt 2>&1
grep "^{" gpurun_out/timing.txt | cut -c1-260
B2F_HARRIS_TILE=64 B2F_HARRIS_TMA=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:harris_fused3 -s 3 -c 1 -o gpurun_out/prof_fused3_64 python tools/harris_timing.py 64 0 > gpurun_out/ncu_fused3_64.log 2>&1
B2F_HARRIS_TILE=48 B2F_HARRIS_TMA=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:harris_fused3 -s 3 -c 1 -o gpurun_out/prof_fused3_48 python tools/harris_timing.py 48 0 > gpurun_out/ncu_fused3_48.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; cut -c1-1500 gpurun_out/bench_r2a.json
