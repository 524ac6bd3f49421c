#!/bin/bash
# round 2, call a: first run of the specialised kernels + new parity tests + bench with and without specialisation
mkdir -p gpurun_out
python -m pytest tests/test_specialize_gpu.py -x -q > gpurun_out/r02a_spec_tests.log 2>&1
echo "spec tests rc=$?"; tail -15 gpurun_out/r02a_spec_tests.log
python -m pytest tests/test_c3_parity_gpu.py tests/test_large_n_gpu.py::test_config4_full_size_vs_oracle -q -s > gpurun_out/r02a_parity_tests.log 2>&1
echo "parity tests rc=$?"; tail -30 gpurun_out/r02a_parity_tests.log
WV_SPECIALIZE=0 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench_interp.json 2> gpurun_out/r02a_bench_interp.err
echo "bench interp rc=$?"; cat gpurun_out/r02a_bench_interp.json | cut -c1-400
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench_spec.json 2> gpurun_out/r02a_bench_spec.err
echo "bench spec rc=$?"; cat gpurun_out/r02a_bench_spec.json | cut -c1-400
tail -5 gpurun_out/r02a_bench_spec.err
