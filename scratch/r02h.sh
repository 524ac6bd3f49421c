#!/bin/bash
mkdir -p gpurun_out
( time python bench.py --steps 3 --warmup 3 > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err ) 2>&1 | grep real
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02h_bench.json'))
print(d['value'], d['e2e'], d['lml_grad_evals_per_sec'])
print(d['roofline'])
print(json.dumps(d['other_configs'], indent=1))
print(d['cpu_baseline'])
PY
tail -5 gpurun_out/r02h_bench.err
