#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_eval_parity_gpu.py tests/test_fit_gpu.py tests/test_specialize_gpu.py -x -q > gpurun_out/r02g_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/r02g_tests.log
python scratch/perf_c3.py 2000 2>&1 | grep -E "per-class|eval 2"
