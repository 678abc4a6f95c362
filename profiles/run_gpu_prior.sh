#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 200 python -m pytest tests/test_gpu_prior.py -m gpu -q --timeout 100 -rf 2>&1 | tail -5 | cut -c1-250
timeout 200 python profiles/prior_timing.py 256 64 1024 2>&1 | grep -v Warning | tee $OUT/prior_timing_r2.txt | cut -c1-700
