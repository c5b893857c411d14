#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_harris_gpu.py tests/test_contour_gpu.py tests/test_config_sizes_gpu.py -q > gpurun_out/pytest_h.log 2>&1; tail -5 gpurun_out/pytest_h.log
for cb in 25165824 52428800 104857600; do
  B2F_CHUNK_BYTES=$cb timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_chunk_$cb.json 2>/dev/null
  python -c "import json;d=json.load(open('gpurun_out/bench_chunk_$cb.json'));print($cb, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'cert_ms', round(d['detail_ms_per_step']['harris_certify_nms_ms'],3))"
done
timeout 600 python bench.py --workload stream8k --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_stream8k_b.json 2>/dev/null
python -c "import json;d=json.load(open('gpurun_out/bench_stream8k_b.json'));print('stream8k value', round(d['value']), 'e2e', round(d['e2e']['value']), d['detail_ms_per_step'])"
