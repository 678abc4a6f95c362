#!/bin/bash
# First GPU call of the next round: validate and A/B the three variants that were written after
# round 1's GPU budget was spent (all default-off).
#   gpurun --timeout 900 -- 'bash profiles/run_gpu_experiments.sh r2'                (one GPU)
#   gpurun --gpus 2 --timeout 900 -- 'bash profiles/run_gpu_experiments.sh r2 2'     (CE reduction)
TAG=${1:-r2}
N=${2:-1}
OUT=gpurun_out
mkdir -p $OUT
bench() {  # name, env...
  name=$1; shift
  if [ "$N" = "1" ]; then
    env "$@" timeout 200 python bench.py --steps 30 --warmup 3 --no-cpu-baseline \
        > $OUT/exp_${TAG}_$name.json 2> $OUT/exp_${TAG}_$name.err
  else
    env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
        --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 3 \
        > $OUT/exp_${TAG}_$name.json 2> $OUT/exp_${TAG}_$name.err
  fi
  python - "$OUT/exp_${TAG}_$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(sys.argv[2], "ms/step %.3f value %.3e p1 %.3f p2 %.3f red %.3f" % (
        d["ms_per_step"], d["value"], r["pass1_ms"], r["pass2_ms"], r["reduce_cov_ms"]))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
if [ "$N" = "1" ]; then
  TB_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_solver.py \
      tests/test_gpu_prior.py -m gpu -q -k experimental --timeout 120 \
      > $OUT/pytest_exp_$TAG.log 2>&1
  tail -5 $OUT/pytest_exp_$TAG.log
  bench shipped X=1
  bench prefetch TB_OPTIONS=prefetch=1
  bench unroll4 TB_OPTIONS=prefetch=2
  bench fusecov TB_FUSE_COV=1
  bench padmap TB_PADMAP=1
  bench padmap_prefetch TB_PADMAP=1 TB_OPTIONS=prefetch=1
  bench fusecov_prefetch TB_FUSE_COV=1 TB_OPTIONS=prefetch=1
else
  bench auto X=1
  bench serial TB_PIPE_CHUNKS=0
  bench ce4 TB_REDUCE=ce TB_PIPE_CHUNKS=4 TB_GRAPH=0
  bench ce8 TB_REDUCE=ce TB_PIPE_CHUNKS=8 TB_GRAPH=0
fi
