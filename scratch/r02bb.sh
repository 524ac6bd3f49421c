#!/bin/bash
for p in 1 2 3; do echo "== WV_FIT_PIECES_PER_STREAM=$p"; WV_FIT_PIECES_PER_STREAM=$p timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.1f e2e %.1f' % (d['value'], d['e2e']['value']))"; done
