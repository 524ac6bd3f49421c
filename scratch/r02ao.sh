#!/bin/bash
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.1f e2e %.1f' % (d['value'], d['e2e']['value']))"
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extras --outcomes 250 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('250 outcomes: value %.1f e2e %.1f' % (d['value'], d['e2e']['value']))"
WV_CHOL_ALL=0 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extras --outcomes 250 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('250 outcomes, WV_CHOL_ALL=0: value %.1f e2e %.1f' % (d['value'], d['e2e']['value']))"
timeout 300 python scratch/search_c2_warm.py 32 2>&1 | tail -2
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
