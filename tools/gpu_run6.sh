#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_contour_gpu.py tests/test_lsd_gpu.py tests/test_rshim_gpu.py tests/test_batch_pipeline_gpu.py tests/test_surf_gpu.py -q > gpurun_out/pytest_new.log 2>&1; tail -15 gpurun_out/pytest_new.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_composite.json 2> gpurun_out/bench_composite.err; cut -c1-2500 gpurun_out/bench_composite.json; tail -3 gpurun_out/bench_composite.err
timeout 900 python bench.py --workload surf --steps 5 --warmup 3 > gpurun_out/bench_surf.json 2> gpurun_out/bench_surf.err; cut -c1-1800 gpurun_out/bench_surf.json; tail -3 gpurun_out/bench_surf.err
timeout 900 python bench.py --workload stream8k --steps 5 --warmup 3 > gpurun_out/bench_stream8k.json 2> gpurun_out/bench_stream8k.err; cut -c1-1800 gpurun_out/bench_stream8k.json; tail -3 gpurun_out/bench_stream8k.err
