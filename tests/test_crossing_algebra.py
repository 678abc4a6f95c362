"""The algebra behind the crossing-list passes, checked on the CPU against
the oracle: while a detector stays in one pixel and one baseline, both LHS passes are LINEAR in
the (Q, U) weights, so a run of n samples can be replaced by ONE record {pixel, baseline, n,
sum Q, sum U} without skipping any sample --

    pass 1   zmap[pix] += a w_d (n cal, sum Q, sum U)
    pass 2   out[amp]  += w_d (n a - (n cal, sum Q, sum U) . m)

This numpy model builds the records from the oracle's pointing exactly as tb_obs_pack_pointing
does (flags are part of the run state) and must reproduce SolverLHS of the reference kernels."""

import numpy as np
import pytest

import helpers as H
from helpers import O, S


def build_records(pb):
    """One record per run of consecutive in-interval samples of a detector with the same local
    pixel (or the same 'flagged' state) and the same baseline."""
    recs = []
    view_off = np.concatenate([[0], np.cumsum(pb.n_amp_views)[:-1]])
    for d in range(pb.n_det):
        cal = pb.weights[d, :, 0]
        for v, iv in enumerate(pb.intervals):
            first, last = int(iv["first"]), int(iv["last"])
            s = np.arange(first, last)
            sm, lp = O.global_to_local(pb.pixels[d, first:last], pb.n_pix_submap, pb.global2local)
            flagged = ((pb.solver_flags[d, first:last] & pb.det_flag_mask) != 0) | (sm < 0)
            key = np.where(flagged, -1, sm * pb.n_pix_submap + lp)
            amp = pb.det_start[d] + view_off[v] + (s - first) // pb.step_length
            brk = np.flatnonzero((np.diff(key) != 0) | (np.diff(amp) != 0)) + 1
            for a, b in zip(np.concatenate([[0], brk]), np.concatenate([brk, [len(s)]])):
                if key[a] < 0:
                    continue  # a run of flagged samples contributes to neither pass
                w = pb.weights[d, first + a:first + b]
                assert np.all(w[:, 0] == cal[first + a])  # the I weight is the constant cal
                recs.append((d, key[a], b - a, amp[a], cal[first + a], w[:, 1].sum(),
                             w[:, 2].sum()))
    return recs


def lhs_from_records(pb, recs, a, covapply):
    zmap = np.zeros((pb.n_local_submap * pb.n_pix_submap, 3))
    for d, pix, n, amp, cal, sq, su in recs:
        if pb.amp_flags[amp] == 0:
            zmap[pix] += a[amp] * pb.det_scale[d] * np.array([n * cal, sq, su])
    covapply(pb.n_local_submap, pb.n_pix_submap, 3, pb.cov.reshape(-1), zmap.reshape(-1))
    out = np.zeros_like(a)
    for d, pix, n, amp, cal, sq, su in recs:
        if pb.amp_flags[amp] == 0:
            m = zmap[pix]
            out[amp] += pb.det_scale[d] * (n * a[amp] - (n * cal * m[0] + sq * m[1] + su * m[2]))
    return out


@pytest.mark.parametrize("name,n_det,n_samp,nside", [("c1", 4, 6000, 64), ("c2", 4, 8000, 64),
                                                     ("c4", 2, 20000, 128)])
def test_crossing_records_reproduce_solver_lhs(name, n_det, n_samp, nside):
    ck = H.checker()
    obs = S.make_observation(name, n_det=n_det, n_samp=n_samp, nside=nside, eps_max=0.03)
    pb = O.build_problem(obs, ck, rcond_threshold=1e-5)
    recs = build_records(pb)
    # nothing is skipped: every unflagged in-interval sample is in exactly one record
    in_view = np.zeros(pb.n_samp, dtype=bool)
    for iv in pb.intervals:
        in_view[iv["first"]:iv["last"]] = True
    n_good = int(np.sum(((pb.solver_flags & pb.det_flag_mask) == 0) & in_view[None, :]))
    assert sum(r[2] for r in recs) == n_good
    assert len(recs) < n_good  # and the list is shorter than the samples
    rng = np.random.default_rng(3)
    a = np.where(pb.amp_flags == 0, rng.standard_normal(pb.n_amp), 0.0)
    ref = O.solver_lhs(pb, ck, a, covapply=ck.cov_apply_diag)
    out = lhs_from_records(pb, recs, a, ck.cov_apply_diag)
    H.assert_close_norm(out, ref, rtol=1e-12, what=f"crossing-record LHS ({name})")
