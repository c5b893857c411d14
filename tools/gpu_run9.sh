#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_harris_gpu.py -q -x > gpurun_out/pytest_harris.log 2>&1; tail -4 gpurun_out/pytest_harris.log
for cfg in "48 0" "64 0"; do timeout 200 python tools/harris_timing.py $cfg; done > gpurun_out/timing.txt 2>&1
grep "^{" gpurun_out/timing.txt | cut -c1-200
timeout 300 python tools/front_timing.py > gpurun_out/front_timing.txt 2>&1; cat gpurun_out/front_timing.txt | tail -3
B2F_HARRIS_TILE=48 timeout 600 ncu --set full --clock-control none --import-source on -k regex:harris_fused3 -s 3 -c 1 -o gpurun_out/prof_r2_fused3_ab python tools/harris_timing.py 48 0 > gpurun_out/ncu_fused3_ab.log 2>&1
