#!/usr/bin/env python
"""Time every stand-alone operator kernel (the drop-in `_libtoast` entry points) on device-resident
buffers of one BASELINE.json workload and report algorithmic GB/s against the measured HBM peak.

    python profiles/bench_kernels.py [--workload c2] [--n-det N] [--reps 5]

Algorithmic bytes per det-sample follow SURVEY.md section 8(d).  Arrays are GBs (>> L2), CUDA
events on the launching stream, best of `reps` after one warm-up."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from toast_b200 import kernels as K  # noqa: E402
from toast_b200 import synthetic as S  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--n-det", type=int, default=None)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    cfg = S.CONFIGS[a.workload]
    n_det = a.n_det or min(cfg["n_det"], 1000)
    n_samp = cfg["n_samp"]
    obs = S.make_observation(a.workload, n_det=n_det, n_samp=n_samp, with_signal=False,
                             flags=False)
    dev = torch.device("cuda")
    nside, nest = obs["nside"], obs["nest"]
    n_submap, nps = S.n_submap_for(nside, 16)
    iv = obs["intervals"]
    n_good = int(sum(int(v["last"] - v["first"]) for v in iv)) * n_det
    idx = np.arange(n_det, dtype=np.int32)
    bore = torch.from_numpy(obs["boresight"]).to(dev)
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    sflags = (torch.rand(n_samp, generator=g, device=dev) < 0.005).to(torch.uint8)
    dflags = (torch.rand((n_det, n_samp), generator=g, device=dev) < 0.01).to(torch.uint8)
    quats = torch.zeros((n_det, n_samp, 4), dtype=torch.float64, device=dev)
    pixels = torch.zeros((n_det, n_samp), dtype=torch.int64, device=dev)
    weights = torch.zeros((n_det, n_samp, 3), dtype=torch.float64, device=dev)
    signal = torch.randn((n_det, n_samp), generator=g, device=dev, dtype=torch.float64)
    hits = np.zeros(n_submap, dtype=np.uint8)
    fp, eps, gam, cal, dw = (obs["focalplane"], obs["epsilon"], obs["gamma"], obs["cal"],
                             obs["detweight"])
    nohwp = None

    K.pointing_fused(fp, bore, sflags, 1, None, None, idx, pixels, None, None, nohwp, iv, hits,
                     nps, nside, nest, eps, gam, cal, False)
    local = np.flatnonzero(hits)
    g2l = np.full(n_submap, -1, dtype=np.int64)
    g2l[local] = np.arange(len(local))
    zmap = torch.zeros((len(local), nps, 3), dtype=torch.float64, device=dev)
    step = obs["step_length"]
    nav = np.array([-(-int(v["last"] - v["first"]) // step) for v in iv], dtype=np.int64)
    per_det = int(nav.sum())
    offs = np.arange(n_det, dtype=np.int64) * per_det
    amps = torch.randn(per_det * n_det, generator=g, device=dev, dtype=torch.float64)
    aflags = torch.zeros(per_det * n_det, dtype=torch.uint8, device=dev)
    hitmap = torch.zeros(len(local) * nps, dtype=torch.int64, device=dev)
    invcov = torch.zeros((len(local), nps, 6), dtype=torch.float64, device=dev)

    cases = [
        ("pointing_detector", 32, lambda: K.pointing_detector(fp, bore, idx, quats, iv, sflags, 1)),
        ("pixels_healpix", 40, lambda: K.pixels_healpix(idx, quats, sflags, 1, idx, pixels, iv,
                                                        hits, nps, nside, nest)),
        ("stokes_weights_IQU", 56, lambda: K.stokes_weights_IQU(idx, quats, idx, weights, nohwp,
                                                                iv, eps, gam, cal, False)),
        ("pointing_fused (boresight -> pixels + weights)", 32,
         lambda: K.pointing_fused(fp, bore, sflags, 1, None, None, idx, pixels, idx, weights,
                                  nohwp, iv, hits, nps, nside, nest, eps, gam, cal, False)),
        ("noise_weight", 16, lambda: K.noise_weight(signal, idx, iv, dw)),
        ("build_noise_weighted", 41,
         lambda: K.build_noise_weighted(g2l, zmap, idx, pixels, idx, weights, idx, signal, idx,
                                        dflags, dw, 1, iv, sflags, 1)),
        ("ops_scan_map_float64 (subtract)", 48,
         lambda: K.ops_scan_map_float64(g2l, nps, zmap, signal, idx, pixels, idx, weights, idx, iv,
                                        1.0, False, True, False)),
        ("template_offset_add_to_signal (batched)", 16,
         lambda: K.template_offset_add_to_signal_batch(step, offs, nav, amps, aflags, idx, signal,
                                                       iv)),
        ("template_offset_project_signal (batched)", 9,
         lambda: K.template_offset_project_signal_batch(idx, signal, idx, dflags, 1, step, offs,
                                                        nav, amps, aflags, iv)),
        ("cov_accum (hits + inverse covariance)", 33,
         lambda: K.cov_accum(g2l, len(local), nps, 3, hitmap, invcov, idx, pixels, idx, weights,
                             idx, dflags, dw, 1, iv, sflags, 1)),
    ]
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    peak = float(peaks.get("hbm_gbs", 6650.0))
    rows = []
    for name, nbytes, fn in cases:
        fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(a.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        gbs = n_good * nbytes / (best * 1e-3) / 1e9
        rows.append(dict(kernel=name, ms=round(best, 3), alg_bytes_per_sample=nbytes,
                         gbs=round(gbs, 1), frac_of_peak=round(gbs / peak, 3),
                         msamp_per_s=round(n_good / best / 1e3, 1)))
    out = dict(workload=a.workload, n_det=n_det, n_samp=n_samp, nside=nside, nest=bool(nest),
               det_samples=n_good, peak_gbs=peak, kernels=rows)
    print(json.dumps(out))
    c2 = rows[3]["ms"] + rows[5]["ms"]
    print(f"# C2 composite (pointing_fused + build_noise_weighted): {c2:.3f} ms, "
          f"{n_good * 41 / (c2 * 1e-3) / 1e9:.0f} GB/s algorithmic at 41 B/sample "
          f"({n_good * 41 / (c2 * 1e-3) / 1e9 / peak:.2f} of peak)", file=sys.stderr)


if __name__ == "__main__":
    main()
