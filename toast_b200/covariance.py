"""covariance_invert / covariance_apply (``covariance.py:20-306``) on the GPU kernels:
batched symmetric 3x3 eigen-inversion with an rcond threshold, and the per-pixel C.m product
(``libtoast/src/toast_map_cov.cpp:246-396, 471-528``)."""

import numpy as np

from . import kernels as KC
from .pixels import PixelData


def covariance_invert(npp, threshold, rcond=None, use_alltoallv=False):
    """In-place inversion of the diagonal-block pixel covariance ``npp`` ([.., .., 6] or
    [.., .., 1]); pixels whose reciprocal condition number is below ``threshold`` are zeroed.
    ``npp`` must already be reduced over processes (``covariance.py:20-133``)."""
    block = npp.n_value
    nnz = int((np.sqrt(1 + 8 * block) - 1) // 2)
    npix = npp.distribution.n_local_submap * npp.distribution.n_pix_submap
    rc = rcond.raw if rcond is not None else np.zeros(npix)
    KC.cov_invert(npix, nnz, npp.raw, rc, float(threshold))


def covariance_apply(npp, m, use_alltoallv=False):
    """m <- npp . m per pixel (``covariance.py:262-306``)."""
    d = npp.distribution
    nnz = m.n_value
    from . import _libtoast as K

    K.cov_apply_diag(d.n_local_submap, d.n_pix_submap, nnz, npp.raw, m.raw)


def covariance_rcond(npp, threshold=1e-8):
    """Reciprocal condition number map of an (uninverted) inverse covariance."""
    tmp = PixelData(npp.distribution, np.float64, n_value=npp.n_value)
    tmp.data[:] = npp.data
    rc = PixelData(npp.distribution, np.float64, n_value=1)
    covariance_invert(tmp, threshold, rcond=rc)
    return rc
