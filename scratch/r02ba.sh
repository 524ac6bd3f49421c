#!/bin/bash
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.1f e2e %.1f' % (d['value'], d['e2e']['value']))"
timeout 300 python scratch/e2e_stages.py 2>&1 | grep -v "Batch(B" | sed -n 12,24p
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
