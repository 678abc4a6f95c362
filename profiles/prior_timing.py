"""Noise-prior kernels on the C4 shard: the bench's noise_prior sub-run alone."""
import json
import os
import sys

sys.path.insert(0, ".")
import torch

import bench
from toast_b200 import lib as L

torch.cuda.set_device(0)
lib = L.load()
peak = 6542.1
if os.path.exists("MEASURED_PEAKS.json"):
    peak = float(json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", peak))
for chunk in [int(x) for x in (sys.argv[1:] or ["256"])]:
    L.check(lib.tb_set_option(b"prior_chunk", chunk))
    r = bench.noise_prior_sub_run(torch.device("cuda", 0), lib, peak)
    r["prior_chunk"] = chunk
    print(json.dumps(r))

# parity of the partitioned solve at each chunk length on a 12-hour view (as
# tests/test_gpu_prior.py::test_partitioned_banded_solve_on_a_12_hour_view)
sys.path.insert(0, "tests")
import numpy as np
from helpers import assert_close_norm
from oracle import offset_prior as OP
from test_offset_prior import build_product, make_case

case = make_case(20, n_amp_views=(43200,), n_det=2)
rng = np.random.default_rng(12)
n = case["n_amp"]
a_in = rng.standard_normal(n)
flags = (rng.random(n) < 0.05).astype(np.uint8)
ref = np.zeros(n)
OP.apply_precond(case["prior"], a_in, flags, ref)
for chunk in [int(x) for x in (sys.argv[1:] or ["256"])]:
    L.check(lib.tb_set_option(b"prior_chunk", chunk))
    prior = build_product(case).finish()
    p_d = torch.full((n,), 7.0, dtype=torch.float64, device="cuda")
    prior.precond(torch.from_numpy(a_in).cuda(), torch.from_numpy(flags).cuda(), p_d)
    err = float(np.max(np.abs(p_d.cpu().numpy() - ref)) / np.max(np.abs(ref)))
    print(json.dumps({"prior_chunk": chunk, "precond_rel_err_vs_scipy_43200": err}))
