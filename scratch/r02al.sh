#!/bin/bash
for m in 0 10000 12000; do echo "== bench WV_CHOL_ALL_MAX=$m"; WV_CHOL_ALL_MAX=$m timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.1f e2e %.1f' % (d['value'], d['e2e']['value']))"; done
for m in 0 10000; do echo "== search WV_CHOL_ALL_MAX=$m"; WV_CHOL_ALL_MAX=$m timeout 300 python scratch/search_c2_warm.py 32 2>&1 | tail -2; done
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
