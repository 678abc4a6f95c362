"""Generate the golden fixtures in tests/golden/ from the REFERENCE's compiled kernels.

Run in the build container (needs /root/reference, compiled by oracle/build_ref.sh):

    python tests/golden/make_golden.py

Inputs are not stored: they are regenerated deterministically by toast_b200.synthetic (same
seed); only the reference OUTPUTS are committed.  The fixtures pin both the C restatement
(oracle/toast_oracle.c, checked on CPU) and the CUDA path (checked on the GPU box, where
/root/reference does not exist).
"""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import toast_oracle as O  # noqa: E402
from toast_b200 import synthetic as S  # noqa: E402
import helpers as H  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# (fixture name, workload, n_det, n_samp)
# (fixture name, workload, n_det, n_samp, nside) -- the ground slices use a coarser nside than
# the full workloads so that a few detectors x 2 minutes still cross-link every pixel
CASES = [
    ("c1_tiny", "c1", 4, 3000, 64),
    ("c2_slice", "c2", 6, 12000, 64),
    ("c5_slice", "c5", 6, 12000, 64),
]


def main():
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    R = O.load_ref()
    if R is None:
        raise SystemExit("oracle/_ref is not built: run oracle/build_ref.sh first")

    # -- healpix primitives at the reference tests' nsides (tests/healpix.py:95-184) -----------
    theta, phi = H.healpix_angle_sets()
    out = {}
    for nside in (1, 256, 16384):
        nest = np.zeros(len(theta), dtype=np.int64)
        ring = np.zeros(len(theta), dtype=np.int64)
        R.healpix_ang2nest(nside, theta, phi, nest)
        R.healpix_ang2ring(nside, theta, phi, ring)
        out[f"nest_{nside}"] = nest
        out[f"ring_{nside}"] = ring
        r2n = np.zeros_like(ring)
        R.healpix_ring2nest(nside, ring, r2n)
        out[f"ring2nest_{nside}"] = r2n
    np.savez_compressed(os.path.join(HERE, "healpix_angles.npz"), **out)

    # -- full pipeline slices -----------------------------------------------------------------
    for name, wl, nd, ns, nside in CASES:
        obs = S.make_observation(wl, n_det=nd, n_samp=ns, eps_max=0.05, nside=nside)
        pb = O.build_problem(obs, R)
        rhs = O.solver_rhs(pb, R, obs["signal"], covapply=R.cov_apply_diag)
        ones = np.where(pb.amp_flags == 0, 1.0, 0.0)
        lhs1 = O.solver_lhs(pb, R, ones, covapply=R.cov_apply_diag)
        amps, hist = O.solve(pb, R, rhs, n_iter_max=12, covapply=R.cov_apply_diag)
        amps2, _ = O.solve(pb, R, rhs, n_iter_max=2, covapply=R.cov_apply_diag)
        idx = np.arange(nd, dtype=np.int32)
        zmap = np.zeros((pb.n_local_submap, pb.n_pix_submap, 3))
        R.build_noise_weighted(pb.global2local, zmap, idx, pb.pixels, idx, pb.weights, idx,
                               obs["signal"], idx, pb.solver_flags, pb.det_scale, 1,
                               pb.intervals, pb.shared_flags, 1, False)
        nz = np.flatnonzero(np.any(zmap.reshape(-1, 3) != 0, axis=1))
        np.savez_compressed(
            os.path.join(HERE, f"{name}.npz"),
            workload=wl, n_det=nd, n_samp=ns, nside=nside,
            pixels=pb.pixels.astype(np.int64),
            weights=pb.weights,
            hit_submaps=pb.hit_submaps,
            zmap_index=nz.astype(np.int64),
            zmap_values=zmap.reshape(-1, 3)[nz],
            rhs=rhs, lhs_of_ones=lhs1, amplitudes=amps, amplitudes_iter2=amps2, history=np.array(hist),
        )
        print(name, "pixels", pb.pixels.shape, "hist", hist[:3], "->", hist[-1])


if __name__ == "__main__":
    main()
