#!/bin/bash
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.1f e2e %.1f' % (d['value'], d['e2e']['value']))"
timeout 600 python -m pytest tests/test_postfit_gpu.py tests/test_fit_gpu.py tests/test_c3_parity_gpu.py tests/test_vgp_gpu.py -x -q 2>&1 | tail -3
