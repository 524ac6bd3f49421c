#!/bin/bash
for B in 750 1000 1500; do for t in 0 1; do echo "== B=$B WV_CHOL_ALL=$t"; WV_CHOL_ALL=$t timeout 200 python scratch/perf_c3.py $B 2>&1 | grep -E "eval 2"; done; done
for B in 250 1000; do for lag in 160 2560; do echo "== B=$B WV_CHOL_ALL=1 lag $lag"; WV_CHOL_LAG=$lag WV_CHOL_ALL=1 timeout 200 python scratch/perf_c3.py $B 2>&1 | grep -E "eval 2"; done; done
for m in 0 5000 7500 10000 15000; do echo "== bench WV_CHOL_ALL_MAX=$m"; WV_CHOL_ALL_MAX=$m timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.1f e2e %.1f' % (d['value'], d['e2e']['value']))"; done
