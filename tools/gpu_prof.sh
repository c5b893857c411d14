#!/bin/bash
# end-of-round profiles: launch list of the composite bench step and a --set full capture of the roofline kernel
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_composite_launches_b16.csv python bench.py --steps 2 --warmup 3 --profile-mode > /dev/null 2>&1
B2F_HARRIS_TILE=48 timeout 500 ncu --set full --clock-control none --import-source on -k regex:harris_fused3 -s 3 -c 1 -o gpurun_out/r2_harris_fused3_final python tools/harris_timing.py 48 0 > gpurun_out/prof_timing.log 2>&1
ls -la gpurun_out/r2_harris_fused3_final.ncu-rep gpurun_out/r2_composite_launches_b16.csv; tail -3 gpurun_out/prof_timing.log
