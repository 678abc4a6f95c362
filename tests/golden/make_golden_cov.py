"""Generate tests/golden/cov_invert.npz: outputs of the REFERENCE's own cov_eigendecompose_diag
(libtoast/src/toast_map_cov.cpp:246-396, LAPACK dsyev) on a fixed set of 3x3 pixel covariances.

The reference is compiled from /root/reference by oracle/build_ref.sh; its LAPACK calls are
forwarded to the OpenBLAS that scipy bundles (oracle/ref_shim/lapack_shim.cpp).  Run in the build
container only:

    python tests/golden/make_golden_cov.py
"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def make_blocks(npix=1500, seed=0):
    """Inverse covariances as CovarianceAndHits accumulates them (sums of w w^T with w = (1, q, u))
    from well conditioned to singular, plus empty pixels."""
    rng = np.random.default_rng(seed)
    ang = rng.uniform(0, np.pi, size=(npix, 8))
    w = np.stack([np.ones_like(ang), np.cos(2 * ang), np.sin(2 * ang)], axis=-1)
    n_obs = rng.integers(1, 9, size=npix)
    w[np.arange(8)[None, :] >= n_obs[:, None]] = 0.0      # 1..8 observations per pixel
    w[:150, 1:, :] = w[:150, :1, :]                        # one orientation only: singular
    spread = rng.uniform(1e-6, 1.0, size=npix)
    w[150:400, 1:, 1:] = w[150:400, :1, 1:] + spread[150:400, None, None] * (
        w[150:400, 1:, 1:] - w[150:400, :1, 1:])           # nearly degenerate orientations
    m = np.einsum("pki,pkj->pij", w, w) * rng.uniform(0.5, 2.0, size=npix)[:, None, None]
    m[400:450] = 0.0                                       # never observed
    iu = np.triu_indices(3)
    return np.ascontiguousarray(m[:, iu[0], iu[1]])


def main():
    from oracle import toast_oracle as O

    ref = O.load_ref()
    assert ref is not None, "build oracle/_ref first (oracle/build_ref.sh)"
    blocks = make_blocks()
    out = dict(blocks=blocks)
    for name, thr in (("1e-3", 1e-3), ("1e-8", 1e-8)):
        d = blocks.reshape(-1).copy()
        rc = np.zeros(len(blocks))
        ref.cov_eigendecompose_diag(1, len(blocks), 3, d, rc, thr, True)
        out[f"inverse_{name}"] = d.reshape(-1, 6)
        out[f"rcond_{name}"] = rc
    d = blocks.reshape(-1).copy()
    rc = np.zeros(len(blocks))
    ref.cov_eigendecompose_diag(1, len(blocks), 3, d, rc, 1e-3, False)
    out["rcond_noinvert"] = rc
    np.testing.assert_array_equal(d.reshape(-1, 6), blocks)  # invert=False leaves the data alone
    np.savez_compressed(os.path.join(HERE, "cov_invert.npz"), **out)
    print({k: v.shape for k, v in out.items()},
          "kept at 1e-3:", int((out["rcond_1e-3"] > 0).sum()),
          "kept at 1e-8:", int((out["rcond_1e-8"] > 0).sum()))


if __name__ == "__main__":
    main()
