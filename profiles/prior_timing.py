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
