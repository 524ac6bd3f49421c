#!/bin/bash
timeout 600 python -m pytest tests/test_kernel_search_gpu.py tests/test_c2_search_parity_gpu.py -x -q 2>&1 | tail -4
timeout 900 python scratch/search_c2_lanes.py 2>&1 | tail -14
