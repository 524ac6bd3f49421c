#!/bin/bash
for t in 0 1; do echo "== WV_FEW_MODELS=$t"; WV_FEW_MODELS=$t timeout 300 python scratch/tail_round.py 2>&1 | grep -v "^ *$" | grep -E "^B=|fit"; done
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python scratch/search_rounds.py 2>&1 | tail -9
