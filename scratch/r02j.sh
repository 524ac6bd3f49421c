#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_elbo_gpu.py tests/test_adam_gpu.py::test_adam_matches_the_numpy_restatement -q -x > gpurun_out/r02j_tests.log 2>&1
echo "rc=$?"; tail -25 gpurun_out/r02j_tests.log
