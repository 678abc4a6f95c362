#!/bin/bash
# round 2, call J: full GPU suite (with the experimental partitioned banded solve), the full bench
# line, one ncu --set full capture of the dominant kernel and the launch list of a short run
OUT=gpurun_out
mkdir -p $OUT
TB_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests -m gpu -q --timeout 200 -rf > $OUT/pytest_r2j.log 2>&1
tail -25 $OUT/pytest_r2j.log | cut -c1-300
timeout 600 python bench.py > $OUT/bench_r2j.json 2> $OUT/bench_r2j.err
tail -c 400 $OUT/bench_r2j.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2j.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("N=1 ms/step", d["ms_per_step"], "value %.3e" % d["value"], "frac", r["frac"], "kernel_ms", r["kernel_ms"])
print("e2e", {k: d["e2e"][k] for k in ("value", "ms_per_step", "seconds_total")})
print("parity", d["parity"])
print("cpu", d.get("cpu_baseline"))
for k, v in d.get("other_workloads", {}).items():
    print(k, {a: v[a] for a in v if a in ("value", "ms_per_step", "path", "error", "add_prior_ms", "apply_precond_ms", "pcg_iteration_with_prior_ms", "pointing_fused_ms", "build_noise_weighted_ms", "frac_of_peak", "add_prior_frac_of_peak", "apply_precond_frac_of_peak", "samples_per_record", "split_block_units")})
PY
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_bx<" -c 1 -f \
  -o $OUT/prof_r2j_fused python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline > $OUT/prof_r2j_fused.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_r2j.csv \
  python bench.py --steps 3 --warmup 1 --no-extras --no-cpu-baseline > $OUT/launches_r2j.log 2>&1
ls -la $OUT/prof_r2j_fused.ncu-rep $OUT/launches_r2j.csv
