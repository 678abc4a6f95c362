#!/bin/bash
# round 2, call B: staged (bulk-async) block-ordered passes: parity + variants (block size, stage
# size, CTAs per SM) on the C4 shard
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_solver.py -m gpu -x -q --timeout 300 -k "variants or chunked or golden or lhs_rhs" > $OUT/pytest_r2c.log 2>&1
tail -8 $OUT/pytest_r2c.log
run() { # name lib fuse
  TB_LIB_PATH=$2 TB_FUSE_LHS=$3 timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline \
      > $OUT/r2c_$1.json 2> $OUT/r2c_$1.err
  python - $OUT/r2c_$1.json "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(sys.argv[2], "ms/step %.3f value %.3e p1 %.3f p2 %.3f red %.3f" % (
        d["ms_per_step"], d["value"], r["pass1_ms"], r["pass2_ms"], r["reduce_cov_ms"]), d["pcg_relative_residuals"][:2])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
    print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
}
D=$PWD/toast_b200
for f in 1 0; do
  run default_f$f $D/libtoastb200.so $f
  for v in s7_c4 s8_c3; do
    run ${v}_f$f $D/libtb_$v.so $f
  done
done
