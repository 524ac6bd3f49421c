#!/bin/bash
SOLO=1 timeout 100 python scratch/perf_c3.py 100 2>&1 | grep -E "eval 2|per-class"
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extras --outcomes 250 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('250 outcomes: value %.1f e2e %.1f' % (d['value'], d['e2e']['value']))"
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('2000 outcomes: value %.1f e2e %.1f' % (d['value'], d['e2e']['value']))"
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
