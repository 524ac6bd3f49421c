#!/bin/bash
for t in 2 0 2; do echo "== WV_CHOL_ALL=$t"; WV_CHOL_ALL=$t timeout 300 python -m pytest tests/test_fit_gpu.py -x -q -k concurrent_sub_batches 2>&1 | grep -E "^E |passed|failed" | head -8; done
