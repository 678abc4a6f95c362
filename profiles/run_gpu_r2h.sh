#!/bin/bash
# round 2, call H: full GPU suite + the new bench line (N=1)
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 > $OUT/pytest_r2h.log 2>&1
tail -15 $OUT/pytest_r2h.log
grep PARITY_REPORT $OUT/pytest_r2h.log | head
timeout 900 python bench.py > $OUT/bench_r2h.json 2> $OUT/bench_r2h.err
tail -c 600 $OUT/bench_r2h.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2h.json").read().strip().splitlines()[-1])
print(json.dumps({k: d[k] for k in d if k not in ("config",)}, indent=1)[:6000])
PY
