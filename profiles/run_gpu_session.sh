#!/bin/bash
# One GPU session: parity tests, bench of the shipped path and its A/B variants, ncu launch list
# and `--set full` captures of the two LHS passes.  Outputs under gpurun_out/ (scratch);
# profiles/summarise.py turns them into the committed summaries.
#   gpurun --timeout 1500 -- 'bash profiles/run_gpu_session.sh <tag>'
TAG=${1:-s3}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $OUT/gpu_$TAG.txt 2>&1

if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1
  echo "pytest exit $?" >> $OUT/pytest_$TAG.log
  tail -5 $OUT/pytest_$TAG.log
fi

timeout 400 python bench.py > $OUT/bench_${TAG}_default.json 2> $OUT/bench_${TAG}_default.err
tail -c 600 $OUT/bench_${TAG}_default.json
for v in ${VARIANTS:-sorted=0 crossings=0 crossings=0,pairw=0}; do
  TB_OPTIONS=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline \
      > $OUT/bench_${TAG}_$v.json 2> $OUT/bench_${TAG}_$v.err
done

if [ "${SKIP_NCU:-0}" != "1" ]; then
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
      --log-file $OUT/launches_$TAG.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline \
      > $OUT/launches_$TAG.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on \
      -k regex:'k_bin_xs|k_lhs_x|k_proj_xs' --launch-skip 6 -c 2 -f -o $OUT/prof_${TAG}_x \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/prof_${TAG}_x.log 2>&1
fi
ls -la $OUT
