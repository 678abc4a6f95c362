#!/bin/bash
# Sweep of the chunk pipeline (chunks x reduction-kernel CTAs per SM) on N GPUs.
#   gpurun --gpus 2 --timeout 900 -- 'bash profiles/run_gpu_pipe_sweep.sh <tag> <N>'
TAG=${1:-sweep}
N=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/sweep_$TAG.txt
CONFIGS=${CONFIGS:-0:12 4:12 8:6 8:3 8:2 4:3 16:3 4:6}
for cfg in $CONFIGS; do
  chunks=${cfg%%:*}; ctas=${cfg##*:}
  TB_MULTIMEM=${TB_MULTIMEM:-0} TB_PIPE_CHUNKS=$chunks TB_OPTIONS=peer_ctas=$ctas timeout 200 \
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 29519 bench.py --gpus $N --steps 20 --warmup 3 \
    > $OUT/sweep_${TAG}_$cfg.json 2> $OUT/sweep_${TAG}_$cfg.err
  python - "$OUT/sweep_${TAG}_$cfg.json" "$cfg" >> $OUT/sweep_$TAG.txt <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(sys.argv[2], "ms/step %.3f value %.3e p1 %.3f p2 %.3f red %.3f" % (
        d["ms_per_step"], d["value"], r["pass1_ms"], r["pass2_ms"], r["reduce_cov_ms"]))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
done
cat $OUT/sweep_$TAG.txt
