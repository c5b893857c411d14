#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_contour_gpu.py tests/test_lsd_gpu.py tests/test_rshim_gpu.py -q -x > gpurun_out/pytest_new.log 2>&1; tail -15 gpurun_out/pytest_new.log
for cfg in "64 0" "48 0" "488 0"; do timeout 200 python tools/harris_timing.py $cfg; done > gpurun_out/timing.txt 2>&1
grep "^{" gpurun_out/timing.txt | cut -c1-200
timeout 600 python -m pytest tests/test_harris_gpu.py -q > gpurun_out/pytest_harris.log 2>&1; tail -4 gpurun_out/pytest_harris.log
timeout 900 python -m pytest tests/test_batch_pipeline_gpu.py tests/test_surf_gpu.py tests/test_canny_gpu.py tests/test_fhog_gpu.py tests/test_otsu_gpu.py -q > gpurun_out/pytest_rest.log 2>&1; tail -5 gpurun_out/pytest_rest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
