#!/bin/bash
# 8 GPUs: multi-rank parity tests at world 2/4/8 + bench
OUT=gpurun_out
TAG=${1:-r2n8}
N=${2:-8}
mkdir -p $OUT
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_peer.py -m gpu -x -q --timeout 600 > $OUT/pytest_$TAG.log 2>&1
tail -6 $OUT/pytest_$TAG.log
bench() {  # name, env...
  name=$1; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
      --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 3 --no-extras \
      > $OUT/${TAG}_$name.json 2> $OUT/${TAG}_$name.err
  python - "$OUT/${TAG}_$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(sys.argv[2], "ms/step %.3f value %.3e p1 %.3f p2 %.3f red %.3f" % (
        d["ms_per_step"], d["value"], r["pass1_ms"], r["pass2_ms"], r["reduce_cov_ms"]),
        "pipeline", r["pipeline"], r["pipeline_tuning_ms"], r["map_reduction"], r["map_reduction_tuning_ms"],
        "parity", d["parity"]["map_reduce_max_rel_err"], "nvlink", d["nvlink"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
    print(open(sys.argv[1].replace(".json", ".err")).read()[-2500:])
PY
}
bench auto X=1
bench serial TB_PIPE_CHUNKS=0
bench p2p TB_MULTIMEM=0
