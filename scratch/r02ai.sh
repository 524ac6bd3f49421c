#!/bin/bash
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02y_bench_n8.json 2> gpurun_out/r02y_bench_n8.err || tail -5 gpurun_out/r02y_bench_n8.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02y_bench_n8.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e'].get('value_gather_true'))
oc=d.get('other_configs',{})
print({k:(v.get('value') or v.get('search_s') or v.get('eval_ms') or v) for k,v in oc.items()})
PY
