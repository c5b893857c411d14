#!/bin/bash
# GPU session: TMA probe under the sanitizer, timings, launch list, Harris tests
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck --print-limit 3 python tools/tma_probe.py 64 > gpurun_out/tma_sanitizer.txt 2>&1
if grep -q "tma probe ok" gpurun_out/tma_sanitizer.txt && grep -q "ERROR SUMMARY: 0 errors" gpurun_out/tma_sanitizer.txt; then
  echo "TMA OK"; TMA_OK=1
  timeout 300 compute-sanitizer --tool memcheck --print-limit 3 python tools/tma_probe.py 108 > gpurun_out/tma_sanitizer108.txt 2>&1; tail -3 gpurun_out/tma_sanitizer108.txt
else
  echo "TMA BROKEN"; TMA_OK=0; head -12 gpurun_out/tma_sanitizer.txt
fi
for cfg in "64 0" "108 0"; do timeout 200 python tools/harris_timing.py $cfg; done > gpurun_out/timing.txt 2>&1
if [ $TMA_OK = 1 ]; then for cfg in "64 1" "108 1"; do timeout 200 python tools/harris_timing.py $cfg; done >> gpurun_out/timing.txt 2>&1; fi
grep "^{" gpurun_out/timing.txt
export B2F_HARRIS_TMA=0
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r2_harris.csv python tools/harris_timing.py 64 0 > /dev/null 2>&1
unset B2F_HARRIS_TMA
if [ $TMA_OK = 1 ]; then
  timeout 900 python -m pytest tests/test_harris_gpu.py -q > gpurun_out/pytest_harris.log 2>&1
else
  B2F_HARRIS_TMA=0 timeout 900 python -m pytest tests/test_harris_gpu.py -q -k "not (64-1 or 108-1)" > gpurun_out/pytest_harris.log 2>&1
fi
tail -15 gpurun_out/pytest_harris.log
