#!/bin/bash
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02y_bench_n2.json 2> gpurun_out/r02y_bench_n2.err || tail -5 gpurun_out/r02y_bench_n2.err
cat gpurun_out/r02y_bench_n2.json | cut -c1-1500
