"""Parity AT THE SIZES THAT ARE BENCHMARKED (BASELINE.json configs C3, C4, C5): the compiled
reference (oracle/_ref) runs a 4-detector subset of the production-size problem -- the full
number of samples, the production nside and the production rcond threshold 1e-8
(ops/mapmaker.py: solve_rcond_threshold) -- and the GPU path must reproduce it on the same
subset: pixels and hit-submaps bit-exact, weights / binned map / RHS / LHS / amplitudes within
1e-10, and the PCG in restart parity (helpers.restart_parity) for every iteration.

The pixel covariance is an INPUT of SolverLHS (mapmaker_solve.py:342-506); it is accumulated on
the device from the whole per-GPU shard (the kernels are held to the reference in
test_gpu_kernels.py / test_gpu_ops.py) and handed to both sides, so that the 4 detectors see the
cross-linked production covariance instead of the singular one they would produce alone.
"""

import json
import os

import numpy as np
import pytest
import torch

import helpers as H
from helpers import O, S, assert_close_norm
from toast_b200 import kernels as K
from toast_b200.solver import DeviceObservation, Destriper

pytestmark = pytest.mark.gpu

RCOND = 1.0e-8   # production solve_rcond_threshold (ops/mapmaker.py)
# (workload, detectors that build the covariance, detectors the reference runs, PCG iterations)
CASES = [("c4", 128, 4, 12), ("c3", 256, 4, 12), ("c5", 256, 4, 12)]


def _shard_covariance(name, n_big, n_sub):
    """Pointing of the first n_big detectors of workload ``name`` at its full length and nside;
    returns the observation dict of the first n_sub detectors (with the flags of the big draw),
    the hit-submap mask, and covariance / rcond at the production threshold (host arrays)."""
    dev = torch.device("cuda")
    big = S.make_observation(name, n_det=n_big, with_signal=False)
    n_samp, nside, nest = big["n_samp"], big["nside"], big["nest"]
    n_submap, nps = S.n_submap_for(nside, 16)
    sflags = torch.from_numpy(big["shared_flags"]).to(dev)
    flags = torch.from_numpy(big["det_flags"]).to(dev)
    flags |= sflags[None, :]
    in_view = torch.zeros(n_samp, dtype=torch.bool, device=dev)
    for v in big["intervals"]:
        in_view[int(v["first"]):int(v["last"])] = True
    flags |= (~in_view).to(torch.uint8)[None, :]
    dobs = DeviceObservation(
        focalplane=big["focalplane"], boresight=big["boresight"], intervals=big["intervals"],
        det_scale=big["detweight"], step_length=big["step_length"], nside=nside, nest=nest,
        n_pix_submap=nps, n_submap=n_submap, global2local=np.zeros(n_submap, dtype=np.int64),
        epsilon=big["epsilon"], gamma=big["gamma"], cal=big["cal"], shared_flags=sflags,
        shared_flag_mask=1, solver_flags=flags, solver_flag_mask=1, compact=False)
    hits = np.zeros(n_submap, dtype=np.uint8)
    dobs.expand_pointing(hits)
    flags |= (dobs.pixels < 0).to(torch.uint8)
    local = np.flatnonzero(hits).astype(np.int64)
    g2l = np.full(n_submap, -1, dtype=np.int64)
    g2l[local] = np.arange(len(local))
    n_loc = len(local)
    idx = np.arange(n_big, dtype=np.int32)
    inv = torch.zeros((n_loc, nps, 6), dtype=torch.float64, device=dev)
    K.cov_accum(g2l, n_loc, nps, 3, None, inv, idx, dobs.pixels, idx, dobs.weights, idx, flags,
                big["detweight"], 1, big["intervals"], None, 0)
    rc = torch.zeros(n_loc * nps, dtype=torch.float64, device=dev)
    K.cov_invert(n_loc * nps, 3, inv, rc, RCOND)
    small = S.make_observation(name, n_det=n_sub)
    small["det_flags"] = np.ascontiguousarray(big["det_flags"][:n_sub])
    assert np.array_equal(small["shared_flags"], big["shared_flags"])
    assert np.array_equal(small["focalplane"], big["focalplane"][:n_sub])
    out = dict(small=small, hits=hits, cov=inv.cpu().numpy(), rcond=rc.cpu().numpy(),
               pixels=dobs.pixels[:n_sub].cpu().numpy(), weights=dobs.weights[:n_sub].cpu().numpy(),
               kept=float((rc > 0).double().mean()))
    del dobs, inv, rc, flags
    torch.cuda.empty_cache()
    return out


@pytest.mark.parametrize("name,n_big,n_sub,n_iter", CASES)
def test_reference_subset_at_production_size(name, n_big, n_sub, n_iter):
    ck = H.checker()
    sc = _shard_covariance(name, n_big, n_sub)
    obs = sc["small"]
    pb = O.build_problem(obs, ck, external=dict(hit_submaps=sc["hits"], cov=sc["cov"],
                                                rcond=sc["rcond"]))
    good = float(np.mean(pb.solver_flags == 0))
    assert good > 0.25, f"{name}: only {good:.2f} of the samples survive the solver flags"
    # pointing of the subset: the reference's, bit for bit
    np.testing.assert_array_equal(sc["pixels"], pb.pixels)
    assert_close_norm(sc["weights"], pb.weights, what="weights")
    assert np.all(sc["hits"][pb.own_hit_submaps != 0] != 0)

    dobs = DeviceObservation(
        focalplane=obs["focalplane"], boresight=obs["boresight"], intervals=obs["intervals"],
        det_scale=pb.det_scale, step_length=pb.step_length, nside=pb.nside, nest=pb.nest,
        n_pix_submap=pb.n_pix_submap, n_submap=pb.n_submap, global2local=pb.global2local,
        epsilon=obs["epsilon"], gamma=obs["gamma"], cal=obs["cal"],
        shared_flags=pb.shared_flags, shared_flag_mask=pb.shared_flag_mask,
        solver_flags=pb.solver_flags, solver_flag_mask=pb.det_flag_mask)
    dobs.expand_pointing(np.zeros(pb.n_submap, dtype=np.uint8))
    ds = Destriper([dobs], pb.n_local_submap, pb.n_pix_submap, pb.cov, pb.offset_var,
                   pb.amp_flags)
    covapply = ck.cov_apply_diag
    sig = torch.from_numpy(obs["signal"]).cuda()
    # the noise-weighted map BEFORE the covariance product is well conditioned: strict 1e-10
    idx = np.arange(pb.n_det, dtype=np.int32)
    zref = np.zeros((pb.n_local_submap, pb.n_pix_submap, 3))
    ck.build_noise_weighted(pb.global2local, zref, idx, pb.pixels, idx, pb.weights, idx,
                            obs["signal"], idx, pb.solver_flags, pb.det_scale, pb.det_flag_mask,
                            pb.intervals, pb.shared_flags, pb.shared_flag_mask, False)
    zgpu = torch.zeros((pb.n_local_submap, pb.n_pix_submap, 3), dtype=torch.float64, device="cuda")
    K.build_noise_weighted(pb.global2local, zgpu, idx, dobs.pixels, idx, dobs.weights, idx, sig,
                           idx, dobs.solver_flags, pb.det_scale, pb.det_flag_mask, pb.intervals,
                           dobs.shared_flags, pb.shared_flag_mask)
    assert_close_norm(zgpu.cpu().numpy(), zref, what="noise-weighted map")
    del zgpu, zref
    # everything behind the covariance product (condition numbers up to 1e8 at the production
    # threshold): 1e-10, or the reference's own summation-order dependence where that is larger
    tol = {}
    binned_ref = O.bin_map(pb, ck, obs["signal"], covapply)
    tol["binned"] = H.order_tolerance(binned_ref, O.bin_map(pb, ck, obs["signal"], covapply,
                                                            reverse=True), rcond=RCOND)
    assert_close_norm(ds.bin_signal([sig]).cpu().numpy(), binned_ref, rtol=tol["binned"][0],
                      what="binned map")
    rhs_ref = O.solver_rhs(pb, ck, obs["signal"], covapply=covapply)
    tol["rhs"] = H.order_tolerance(rhs_ref, O.solver_rhs(pb, ck, obs["signal"], covapply=covapply,
                                                         reverse=True), rcond=RCOND)
    assert_close_norm(ds.rhs([sig]).cpu().numpy(), rhs_ref, rtol=tol["rhs"][0], what="RHS")
    rng = np.random.default_rng(11)
    a = np.where(pb.amp_flags == 0, rng.standard_normal(pb.n_amp), 0.0)
    lhs_ref = O.solver_lhs(pb, ck, a, covapply=covapply)
    tol["lhs"] = H.order_tolerance(lhs_ref, O.solver_lhs(pb, ck, a, covapply=covapply,
                                                         reverse=True), rcond=RCOND)
    q = torch.zeros(pb.n_amp, dtype=torch.float64, device="cuda")
    ds.lhs(torch.from_numpy(a).cuda(), q)
    assert_close_norm(q.cpu().numpy(), lhs_ref, rtol=tol["lhs"][0], what="LHS")
    if ds._blocked():   # and the separate-launch form of the same passes
        ds.fuse_lhs = False
        q2 = torch.zeros_like(q)
        ds.lhs(torch.from_numpy(a).cuda(), q2)
        assert_close_norm(q2.cpu().numpy(), lhs_ref, rtol=tol["lhs"][0],
                          what="LHS (pass 1 / covariance / pass 2)")
        ds.fuse_lhs = True

    # PCG: the reference's own iteration states, one device iteration from each
    trace = []
    amps_ref, hist_ref = O.solve(pb, ck, rhs_ref, n_iter_max=n_iter, covapply=covapply,
                                 trace=trace)
    worst = H.restart_parity(ds, pb, trace, rtol=tol["lhs"][0], what=name)
    amps3_ref = trace[3]["x"] if len(trace) > 3 else amps_ref
    amps3, _ = ds.solve(torch.from_numpy(rhs_ref).cuda(), n_iter_max=min(3, len(trace)))
    assert_close_norm(amps3.cpu().numpy(), amps3_ref, rtol=8 * tol["lhs"][0],
                      what="amplitudes after 3 iterations")
    # free-running history: where does it leave 1e-10 (reported)
    _, hist = ds.solve(torch.from_numpy(rhs_ref).cuda(), n_iter_max=n_iter)
    first, dev = H.first_iteration_over(hist, hist_ref)
    assert dev[0] <= 8 * tol["lhs"][0]
    report = dict(tolerances={k: dict(used=v[0], reference_order_dependence=v[1])
                              for k, v in tol.items()},
                  workload=name, detectors=n_sub, covariance_detectors=n_big,
                  n_samp=int(pb.n_samp), nside=int(pb.nside), unflagged_fraction=good,
                  pixels_kept_by_rcond=sc["kept"], restart_parity_worst=max(worst),
                  free_running_first_iteration_over_1e10=first,
                  free_running_deviation=[float(x) for x in dev])
    print("PARITY_REPORT " + json.dumps(report))
    out = os.path.join(H.ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, f"parity_fullsize_{name}.json"), "w") as f:
            json.dump(report, f)
