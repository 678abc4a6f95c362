#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
run() { # name lib fuse
  TB_LIB_PATH=$2 TB_FUSE_LHS=$3 timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline \
      > $OUT/r2g_$1.json 2> $OUT/r2g_$1.err
  python - $OUT/r2g_$1.json "$1" <<'PY'
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
  for v in s8_c4_p0 s7_c4_p1 s7_c4_p0 s8_c3_p1 s6_c4_p0; do
    run ${v}_f$f $D/libtb_$v.so $f
  done
done
