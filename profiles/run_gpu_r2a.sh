#!/bin/bash
# round 2, call A: validate the block-ordered passes (tb_blocked.cu) and A/B them against the
# pixel-sorted passes of round 1 on the C4 shard.
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_solver.py -m gpu -x -q --timeout 300 > $OUT/pytest_r2a.log 2>&1
tail -15 $OUT/pytest_r2a.log
for v in "blocked=1" "blocked=0"; do
  for f in 1 0; do
    TB_OPTIONS=$v TB_FUSE_LHS=$f timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline \
      > $OUT/r2a_${v}_f$f.json 2> $OUT/r2a_${v}_f$f.err
    python - $OUT/r2a_${v}_f$f.json "$v fuse=$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(sys.argv[2], "ms/step %.3f value %.3e p1 %.3f p2 %.3f red %.3f e2e %.3f" % (
        d["ms_per_step"], d["value"], r["pass1_ms"], r["pass2_ms"], r["reduce_cov_ms"], d["e2e"]["ms_per_step"]), d["pcg_relative_residuals"][:3])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
    print(open(sys.argv[1].replace(".json", ".err")).read()[-2000:])
PY
  done
done
