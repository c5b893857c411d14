#!/bin/bash
# one validation pass on the GPU box: tests, smoke, the three bench lines, the pipeline traces, SURF profiles
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/val_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/val_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/val_smoke.log
timeout 600 python bench.py > gpurun_out/val_bench_composite.log 2>&1
timeout 600 python bench.py --workload surf > gpurun_out/val_bench_surf.log 2>&1
timeout 600 python bench.py --workload stream8k > gpurun_out/val_bench_stream8k.log 2>&1
timeout 300 python tools/e2e_trace.py > gpurun_out/val_e2e_trace.log 2>&1
timeout 300 python tools/e2e_probe.py > gpurun_out/val_e2e_probe.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_surf_launches_b16.csv python bench.py --workload surf --steps 2 --warmup 3 --profile-mode > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:surf_pyramid0 -s 2 -c 1 -o gpurun_out/r2_surf_pyramid0 python bench.py --workload surf --steps 1 --warmup 1 --profile-mode > /dev/null 2>&1
tail -3 gpurun_out/val_tests.log; tail -2 gpurun_out/val_smoke.log; for w in composite surf stream8k; do tail -1 gpurun_out/val_bench_$w.log | cut -c1-220; done; cat gpurun_out/val_e2e_trace.log gpurun_out/val_e2e_probe.log
