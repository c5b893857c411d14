#!/bin/bash
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --print-limit 5 python tools/sanitize_smoke.py > gpurun_out/sanitize_r2.log 2>&1; grep -E "^ok|ERROR SUMMARY|Invalid|Error" gpurun_out/sanitize_r2.log | head -20
timeout 600 python -m pytest tests/test_rshim_gpu.py tests/test_batch_pipeline_gpu.py -q > gpurun_out/pytest_shim.log 2>&1; tail -4 gpurun_out/pytest_shim.log
# launch list of the composite step + full captures of the kernels the round-1 profiles lacked
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2_composite.csv python bench.py --steps 2 --warmup 3 --profile-mode > gpurun_out/ncu_launch.log 2>&1
for k in canny_grad_nms_spec2 hyst_local fhog_pixel8 nms_tolerant; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/prof_r2_$k python bench.py --steps 1 --warmup 3 --profile-mode > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -8
