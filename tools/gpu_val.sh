#!/bin/bash
# one validation pass on the GPU box: tests, smoke, the e2e probe, the three bench lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/val_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/val_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/val_smoke.log
timeout 300 python tools/e2e_probe.py > gpurun_out/val_probe.log 2>&1
timeout 600 python bench.py > gpurun_out/val_bench.log 2>&1
tail -3 gpurun_out/val_tests.log; tail -2 gpurun_out/val_smoke.log; cat gpurun_out/val_probe.log; tail -1 gpurun_out/val_bench.log
