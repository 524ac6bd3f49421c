#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_adam_gpu.py::test_adam_matches_the_numpy_restatement tests/test_penalization_gpu.py -q -s > gpurun_out/r02i_tests.log 2>&1
echo "adam/pen rc=$?"; tail -15 gpurun_out/r02i_tests.log
python -m pytest tests/test_c2_search_parity_gpu.py -q -s > gpurun_out/r02i_c2.log 2>&1
echo "c2 rc=$?"; tail -12 gpurun_out/r02i_c2.log
