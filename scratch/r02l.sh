#!/bin/bash
mkdir -p gpurun_out
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r02l_bench_n2.json 2> gpurun_out/r02l_bench_n2.err ) 2>&1 | grep real
echo "rc=$?"; tail -3 gpurun_out/r02l_bench_n2.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02l_bench_n2.json'))
print(d['value'], d['e2e'], d['n_gpus'])
print(json.dumps(d['other_configs'], indent=1)[:3000])
PY
