#!/bin/bash
# ncu launch list (gpu__time_duration only) of a short bench run: per-kernel share of the step
OUT=gpurun_out
TAG=${1:-r2}
mkdir -p $OUT
for f in 1 0; do
TB_FUSE_LHS=$f timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $OUT/launches_${TAG}_f$f.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline > $OUT/launches_${TAG}_f$f.log 2>&1
python - $OUT/launches_${TAG}_f$f.csv <<'PY'
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; iN = hdr.index("Kernel Name"); iV = hdr.index("Metric Value")
seq = [(r[iN].split("(")[0][-60:], float(r[iV].replace(",", ""))) for r in rows[1:]]
print("last 45 launches (us):")
for n, v in seq[-45:]:
    print("  %-60s %9.1f" % (n, v / 1000 if v > 10000 else v))
PY
done
