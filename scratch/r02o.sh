#!/bin/bash
python -m pytest tests/test_postfit_gpu.py tests/test_fit_gpu.py tests/test_vgp_gpu.py -x -q 2>&1 | tail -3
python scratch/e2e_stages.py > gpurun_out/r02o_e2e_stages.log 2>&1; grep -A12 "^rep 2" gpurun_out/r02o_e2e_stages.log
