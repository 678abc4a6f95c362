"""The driver-facing contract of bench.py that can be checked without a GPU: the reference arm
prints ONE JSON line with the agreed keys, non-zero ranks of a torchrun launch stay silent, and
the GPU arm refuses to run without a CUDA device (there is no CPU fallback)."""

import json
import os
import subprocess
import sys

import helpers as H

BENCH = os.path.join(H.ROOT, "bench.py")


def _run(args, env=None, timeout=300):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, env=e,
                          timeout=timeout)


def test_reference_arm_prints_one_json_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["metric"] == "detector-samples/s per destriper PCG iteration"
    assert d["unit"] == "det-samples/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}


def test_reference_arm_is_silent_on_other_ranks():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
             env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_a_device():
    import torch

    if torch.cuda.is_available():
        return  # on the GPU box the arm runs; nothing to check here
    r = _run(["--steps", "1", "--warmup", "0"])
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)
