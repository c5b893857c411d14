#!/bin/bash
mkdir -p gpurun_out
for w in composite surf; do
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --workload $w --no-cpu-baseline > gpurun_out/mgpu2_$w.log 2>&1; echo "rc=$?" >> gpurun_out/mgpu2_$w.log
tail -2 gpurun_out/mgpu2_$w.log | cut -c1-1500
done
timeout 300 python -m pytest tests -m gpu -q -x -k "two_devices or second_device or multi_device" 2>&1 | tail -3
