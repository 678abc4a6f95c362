#!/bin/bash
# Two-GPU validation: the 2-rank parity tests and bench.py with the three map-reduction modes.
#   gpurun --gpus 2 --timeout 900 -- 'bash profiles/run_gpu_n2.sh <tag>'
TAG=${1:-n2}
N=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_peer.py tests/test_gpu_solver.py -m gpu -x -q \
    > $OUT/pytest_$TAG.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_$TAG.log
tail -5 $OUT/pytest_$TAG.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
      --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 \
      > $OUT/bench_${TAG}_$name.json 2> $OUT/bench_${TAG}_$name.err
  tail -c 300 $OUT/bench_${TAG}_$name.err
}
run auto X=1
run p2p TB_MULTIMEM=0
run nccl TB_FUSED_REDUCE=0
ls -la $OUT
