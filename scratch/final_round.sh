#!/bin/bash
# Round-end GPU pass: the whole GPU suite, smoke(), the default bench line and the reference arm, then the ncu launch
# list of one bench step.  Usage: bash scratch/final_round.sh <tag>
TAG=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/${TAG}_gpu_tests.log 2>&1
tail -5 gpurun_out/${TAG}_gpu_tests.log
timeout 120 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
timeout 300 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || tail -5 gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_bench.json
timeout 200 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null
cat gpurun_out/${TAG}_bench_reference.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
wc -l gpurun_out/${TAG}_launches_bench.csv
