#!/bin/bash
# 2 GPUs: copy-engine form of the chunk reduction vs the SM kernel
OUT=gpurun_out
TAG=${1:-r2n2b}
mkdir -p $OUT
bench() {  # name, env...
  name=$1; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
      --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 3 --no-extras \
      > $OUT/${TAG}_$name.json 2> $OUT/${TAG}_$name.err
  python - "$OUT/${TAG}_$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(sys.argv[2], "ms/step %.3f value %.3e p1 %.3f p2 %.3f red %.3f" % (
        d["ms_per_step"], d["value"], r["pass1_ms"], r["pass2_ms"], r["reduce_cov_ms"]),
        "pipeline", r["pipeline"], r["pipeline_tuning_ms"], "residuals", d["pcg_relative_residuals"][:2])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
    print(open(sys.argv[1].replace(".json", ".err")).read()[-2500:])
PY
}
bench ce4_graph TB_REDUCE=ce TB_PIPE_CHUNKS=4 TB_GRAPH=1
bench ce3_graph TB_REDUCE=ce TB_PIPE_CHUNKS=3 TB_GRAPH=1
bench ce6_graph TB_REDUCE=ce TB_PIPE_CHUNKS=6 TB_GRAPH=1
bench sm3 TB_PIPE_CHUNKS=3
bench sm5 TB_PIPE_CHUNKS=5
