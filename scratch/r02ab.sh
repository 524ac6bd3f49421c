#!/bin/bash
timeout 900 python -m pytest tests/test_fit_gpu.py tests/test_kernel_search_gpu.py tests/test_c2_search_parity_gpu.py -x -q 2>&1 | tail -4
timeout 600 python scratch/search_c2_warm.py 0 4 8 16 32 64 2>&1 | tail -14
