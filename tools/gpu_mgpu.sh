#!/bin/bash
# multi-GPU sanity: the bench lines of the composite and SURF workloads under torchrun (N = $1, default 2)
N=${1:-2}
mkdir -p gpurun_out
for w in composite surf; do
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload $w --no-cpu-baseline > gpurun_out/mgpu${N}_$w.log 2>&1; echo "rc=$?" >> gpurun_out/mgpu${N}_$w.log
tail -2 gpurun_out/mgpu${N}_$w.log | python -c "
import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(d['config']['workload'], 'N', d['n_gpus'], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'clocks', d['clocks']['sm_mhz'])
    else: print(l)"
done
