#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_c3_parity_gpu.py -q -s > gpurun_out/r02f_c3_parity.log 2>&1
echo "c3 parity rc=$?"; grep -E "structure_identical|objective rel|status_agree|passed|failed|Error" gpurun_out/r02f_c3_parity.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02f_bench.json'))
print(d['value'], d['e2e']['value'], d['lml_grad_evals_per_sec'], d['roofline_groups']['class_ms'], d['fit_status_hist'])
PY
