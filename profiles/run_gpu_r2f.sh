#!/bin/bash
# round 2, call D: warp-tile block-ordered passes with direct REDs: parity, bench, ncu
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_solver.py -m gpu -x -q --timeout 300 -k "variants or chunked or golden or lhs_rhs" > $OUT/pytest_r2f.log 2>&1
tail -4 $OUT/pytest_r2f.log
for f in 1 0; do
  TB_FUSE_LHS=$f timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > $OUT/r2f_f$f.json 2> $OUT/r2f_f$f.err
  python - $OUT/r2f_f$f.json "fuse=$f" <<'PY'
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
done
bash profiles/run_gpu_ncu_bx.sh r2f_bx > /dev/null 2>&1
ls -la $OUT/prof_r2f*
