#!/bin/bash
for s in 2 4 6 8; do
  echo "== WV_FIT_STREAMS=$s"
  WV_FIT_STREAMS=$s timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.1f e2e %.1f' % (d['value'], d['e2e']['value']))"
done
