#!/bin/bash
for t in 0 1; do echo "== WV_SPEC_GRAD_PREFETCH=$t"; WV_SPEC_GRAD_PREFETCH=$t timeout 300 python scratch/perf_c3.py 2000 2>&1 | grep -E "per-class|eval 2"; done
