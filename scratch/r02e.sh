#!/bin/bash
mkdir -p gpurun_out
for v in "4 3" "4 2" "3 2" "5 2"; do
  set -- $v
  echo "== gram minb $1 grad minb $2"
  WV_SPEC_GRAM_MINB=$1 WV_SPEC_GRAD_MINB=$2 WV_RTC_CACHE=/tmp/rtc_$1_$2 python scratch/perf_c3.py 2000 2>&1 | grep per-class
done
