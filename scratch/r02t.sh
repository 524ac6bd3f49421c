#!/bin/bash
# per-class DRAM traffic of ONE config-3 evaluation of 2000 models (all launches), and the launch list of bench.py
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02t_eval_dram.csv python scratch/perf_c3.py 2000 > gpurun_out/r02t_ncu.log 2>&1
tail -3 gpurun_out/r02t_ncu.log
wc -l gpurun_out/r02t_eval_dram.csv
