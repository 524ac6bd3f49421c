#!/bin/bash
for v in "0 4" "0 3" "1 4"; do
  set -- $v
  echo "== cmask $1 gram minb $2"
  WV_SPEC_CMASK=$1 WV_SPEC_GRAM_MINB=$2 WV_RTC_CACHE=/tmp/rtc_$1_$2 python scratch/perf_c3.py 2000 2>&1 | grep per-class
done
python -m pytest tests/test_specialize_gpu.py -x -q 2>&1 | tail -3
