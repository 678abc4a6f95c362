#!/bin/bash
# round 2, final call: smoke, the full GPU suite, the full bench line, the ncu capture the bench's
# `traffic` field points at, and the launch list of a short run
OUT=gpurun_out
TAG=${1:-r2final}
mkdir -p $OUT
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
TB_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests -m gpu -q --timeout 200 -rf > $OUT/pytest_$TAG.log 2>&1
tail -8 $OUT/pytest_$TAG.log | cut -c1-300
TB_E2E_LHS=1 timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
tail -c 300 $OUT/bench_$TAG.err
python - $OUT/bench_$TAG.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d["roofline"]
print("N=1 ms/step", d["ms_per_step"], "value %.3e" % d["value"], "frac", r["frac"], "kernel_ms", r["kernel_ms"], "traffic", r["traffic"], r["traffic_note"])
print("e2e", {k: d["e2e"][k] for k in ("value", "ms_per_step", "seconds_total")}, "e2e_lhs", d.get("e2e_solver_lhs"))
print("parity", {k: v for k, v in d["parity"].items() if "note" not in k and "check" not in k})
for k, v in d.get("other_workloads", {}).items():
    print(k, {a: v[a] for a in v if a in ("value", "ms_per_step", "error", "add_prior_ms", "apply_precond_ms", "pcg_iteration_with_prior_ms", "add_prior_frac_of_peak", "apply_precond_frac_of_peak", "prior_build_host_s")})
PY
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_bx<" -c 1 -f \
  -o $OUT/prof_${TAG}_fused python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline > $OUT/prof_${TAG}_fused.log 2>&1
ls -la $OUT/prof_${TAG}_fused.ncu-rep
