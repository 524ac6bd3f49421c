#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests/ -q -m gpu -x > gpurun_out/r02k_gpu_tests.log 2>&1 ) 2>&1 | grep real
echo "rc=$?"; tail -30 gpurun_out/r02k_gpu_tests.log
