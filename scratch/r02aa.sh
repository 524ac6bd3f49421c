#!/bin/bash
timeout 900 python -m pytest tests/test_fit_gpu.py tests/test_kernel_search_gpu.py tests/test_c2_search_parity_gpu.py tests/test_lbfgs_warp_gpu.py -x -q 2>&1 | tail -8
for t in 0 8 16 32 64; do echo "== WV_SEARCH_TAIL=$t"; WV_SEARCH_TAIL=$t timeout 300 python scratch/search_c2.py 200 5 2>&1 | grep "config 2"; done
