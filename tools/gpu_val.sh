#!/bin/bash
# one validation pass on the GPU box: tests, smoke, the three bench lines, memcheck, the pipeline traces
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/val_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/val_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/val_smoke.log
timeout 600 python bench.py > gpurun_out/val_bench_composite.log 2>&1
timeout 600 python bench.py --workload surf > gpurun_out/val_bench_surf.log 2>&1
timeout 600 python bench.py --workload stream8k > gpurun_out/val_bench_stream8k.log 2>&1
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/val_bench_reference.log 2>&1
timeout 300 python tools/e2e_trace.py > gpurun_out/val_e2e_trace.log 2>&1
timeout 300 python tools/e2e_trace.py 8k > gpurun_out/val_e2e_trace_8k.log 2>&1
timeout 300 python tools/e2e_probe.py > gpurun_out/val_e2e_probe.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py > gpurun_out/val_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/val_memcheck.log
tail -3 gpurun_out/val_tests.log; tail -2 gpurun_out/val_smoke.log; for w in composite surf stream8k reference; do tail -1 gpurun_out/val_bench_$w.log | cut -c1-200; done; tail -4 gpurun_out/val_memcheck.log; tail -3 gpurun_out/val_e2e_probe.log
