#!/bin/bash
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02zz_bench_n$N.json 2> gpurun_out/r02zz_bench_n$N.err || tail -5 gpurun_out/r02zz_bench_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/r02zz_bench_n$N.json'))
print('N=$N value', d['value'], 'e2e', d['e2e']['value'], d['e2e'].get('value_gather_true'))
oc=d.get('other_configs',{})
print({k:(v.get('value') or v.get('search_s') or v.get('eval_ms') or v) for k,v in oc.items()})
PY
