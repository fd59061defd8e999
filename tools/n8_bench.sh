#!/bin/bash
# 8-GPU run of the driver's own command line (bench.py under torchrun) + the reference arm.
mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"; tail -c 300 gpurun_out/bench_n$N.err
python - <<P
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',json.dumps(d['e2e']))
for e in d['extra']:
    e.pop('kernels',None); print(json.dumps(e)[:900])
P
