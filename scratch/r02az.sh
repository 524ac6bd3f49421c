#!/bin/bash
for o in 500 1000; do for s in 1 2 4; do
  echo "== outcomes $o WV_FIT_STREAMS=$s"
  WV_FIT_STREAMS=$s timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extras --outcomes $o 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.1f e2e %.1f' % (d['value'], d['e2e']['value']))"
done; done
