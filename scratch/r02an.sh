#!/bin/bash
for rep in 1 2; do
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | grep -E "^E  |passed|failed|^tests/.*Error" | head -12
done
