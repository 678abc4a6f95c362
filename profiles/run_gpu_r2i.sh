#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python profiles/e2e_stages.py 2>&1 | tail -6
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 > $OUT/pytest_r2i.log 2>&1
tail -15 $OUT/pytest_r2i.log
grep PARITY_REPORT $OUT/pytest_r2i.log | cut -c1-600 | head
