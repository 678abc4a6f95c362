"""world_size-2 gloo tests (CPU) of the multi-rank host logic: detector sharding, the map
all-reduce (PixelData.sync_allreduce), the hit-submap union and the amplitude dot products.
The per-sample compute is stood in for by the oracle so the test runs without a GPU; on the
GPU the same collectives run over NCCL on device buffers (toast_b200.solver.Destriper)."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import helpers as H  # noqa: F401
from helpers import O, S


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_det_total, n_samp, out):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "1"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from toast_b200.data import Comm
        from toast_b200.pixels import PixelData, PixelDistribution
        from toast_b200.templates import Amplitudes

        comm = Comm()
        assert (comm.world_rank, comm.world_size) == (rank, world)
        n_det = n_det_total // world
        obs = S.make_observation("c1", n_det=n_det, n_samp=n_samp, det_first=rank * n_det,
                                 nside=32)
        pb = O.Problem()
        pb.n_det, pb.n_samp, pb.nside, pb.nest = n_det, n_samp, 32, True
        pb.n_submap, pb.n_pix_submap = S.n_submap_for(32, 16)
        pb.focalplane, pb.boresight, pb.intervals = (obs["focalplane"], obs["boresight"],
                                                     obs["intervals"])
        pb.epsilon, pb.gamma, pb.cal, pb.IAU = obs["epsilon"], obs["gamma"], obs["cal"], False
        pb.hwp, pb.shared_flags, pb.shared_flag_mask = np.zeros(1), obs["shared_flags"], 1
        pixels, weights, hits = O.expand_pointing(pb, O)
        # union of hit submaps over ranks -> identical pixel distribution everywhere
        comm.allreduce_(hits, op="max")
        local = np.flatnonzero(hits)
        dist_ = PixelDistribution(12 * 32 * 32, pb.n_submap, local, comm=comm.comm_world)
        z = PixelData(dist_, np.float64, n_value=3)
        idx = np.arange(n_det, dtype=np.int32)
        O.build_noise_weighted(dist_.global_submap_to_local, z.data, idx, pixels, idx, weights,
                               idx, obs["signal"], idx, obs["det_flags"], obs["detweight"], 1,
                               obs["intervals"], obs["shared_flags"], 1)
        z.sync_allreduce()
        a = Amplitudes(comm, 10 * world, 10)
        a.local[:] = rank + 1.0
        a.local_flags[0] = 1
        dot = a.dot(a)
        if rank == 0:
            out.put((hits.copy(), z.data.copy(), dot))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_binning_equals_single_rank():
    world, n_det_total, n_samp = 2, 4, 3000
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_det_total, n_samp, out))
             for r in range(world)]
    for p in procs:
        p.start()
    hits, zmap, dot = out.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    # single-rank reference of the same 4-detector job
    obs = S.make_observation("c1", n_det=n_det_total, n_samp=n_samp, nside=32)
    pb = O.Problem()
    pb.n_det, pb.n_samp, pb.nside, pb.nest = n_det_total, n_samp, 32, True
    pb.n_submap, pb.n_pix_submap = S.n_submap_for(32, 16)
    pb.focalplane, pb.boresight, pb.intervals = obs["focalplane"], obs["boresight"], obs["intervals"]
    pb.epsilon, pb.gamma, pb.cal, pb.IAU = obs["epsilon"], obs["gamma"], obs["cal"], False
    pb.hwp, pb.shared_flags, pb.shared_flag_mask = np.zeros(1), obs["shared_flags"], 1
    pixels, weights, hits1 = O.expand_pointing(pb, O)
    np.testing.assert_array_equal(hits, hits1)
    local, g2l = O.pixel_distribution(hits1)
    z1 = np.zeros((len(local), pb.n_pix_submap, 3))
    idx = np.arange(n_det_total, dtype=np.int32)
    # the sharded job draws each rank's timestream from its own RNG stream
    sig = np.vstack([S.make_observation("c1", n_det=2, n_samp=n_samp, det_first=2 * r,
                                        nside=32)["signal"] for r in range(2)])
    dfl = np.vstack([S.make_observation("c1", n_det=2, n_samp=n_samp, det_first=2 * r,
                                        nside=32)["det_flags"] for r in range(2)])
    sfl = S.make_observation("c1", n_det=2, n_samp=n_samp, det_first=0, nside=32)["shared_flags"]
    # shared flags differ per rank stream in the synthetic generator; bin rank by rank instead
    z1[:] = 0
    for r in range(2):
        o = S.make_observation("c1", n_det=2, n_samp=n_samp, det_first=2 * r, nside=32)
        pbr = O.Problem()
        pbr.__dict__.update(pb.__dict__)
        pbr.n_det = 2
        pbr.focalplane, pbr.epsilon, pbr.gamma, pbr.cal = (o["focalplane"], o["epsilon"],
                                                           o["gamma"], o["cal"])
        pbr.shared_flags = o["shared_flags"]
        px, wt, _ = O.expand_pointing(pbr, O)
        i2 = np.arange(2, dtype=np.int32)
        O.build_noise_weighted(g2l, z1, i2, px, i2, wt, i2, o["signal"], i2, o["det_flags"],
                               o["detweight"], 1, o["intervals"], o["shared_flags"], 1)
    assert np.allclose(zmap, z1, rtol=1e-13, atol=1e-13 * np.abs(z1).max())
    assert dot == 9 * 1.0 + 9 * 4.0
    del sig, dfl, sfl


# ------------------------------------------------------------------------------------------------
# ops.MapMaker, detector-sharded over two ranks: its multi-rank host logic (union of the hit
# submaps, hit map / inverse covariance summed over the ranks, global amplitude count, rank-local
# amplitude slices) with the oracle standing in for the device side (tests/fake_device.py: the map
# and the dot products are summed over the ranks exactly where the device solver does it)
# ------------------------------------------------------------------------------------------------
def _shard(obs, first, count):
    """Detectors [first, first + count) of a synthetic observation (shared arrays kept)."""
    sl = slice(first, first + count)
    out = dict(obs)
    out["n_det"] = count
    for key in ("focalplane", "epsilon", "gamma", "cal", "detweight", "sigma", "det_flags",
                "signal"):
        if key in obs:
            out[key] = np.ascontiguousarray(obs[key][sl])
    return out


def _mapmaker_worker(rank, world, port, n_det_total, n_samp, out):
    import pytest as _pytest
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "1"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mp_ = _pytest.MonkeyPatch()
    try:
        import fake_device
        from toast_b200 import ops
        from toast_b200.data import Data, observation_from_synthetic
        from toast_b200.templates import Offset

        fake_device.install(mp_)
        full = S.make_observation("c1", n_det=n_det_total, n_samp=n_samp, nside=32, eps_max=0.03)
        per = n_det_total // world
        obs = _shard(full, rank * per, per)
        data = Data()
        assert data.comm.world_size == world
        data.obs.append(observation_from_synthetic(obs))
        dp = ops.PointingDetectorSimple(view="scanning", shared_flags="flags", shared_flag_mask=1)
        pix = ops.PixelsHealpix(detector_pointing=dp, nside=32, nest=obs["nest"],
                                create_dist="pixel_dist")
        wts = ops.StokesWeights(detector_pointing=dp, mode="IQU")
        binning = ops.BinMap(pixel_dist="pixel_dist", covariance="cov", pixel_pointing=pix,
                             stokes_weights=wts, noise_model="noise_model", full_pointing=True)
        tmpl = Offset(name="baselines", step_time=obs["step_time"], times="times",
                      noise_model="noise_model")
        tmat = ops.TemplateMatrix(templates=[tmpl], amplitudes="amplitudes")
        mapper = ops.MapMaker(name="mm", det_data="signal", binning=binning, template_matrix=tmat,
                              solve_rcond_threshold=1.0e-3, map_rcond_threshold=1.0e-3,
                              iter_max=4, iter_min=4, convergence=1.0e-30, device="cpu")
        mapper.apply(data)
        amps = data["amplitudes"]["baselines"]
        res = dict(hits=data["mm_hits"].raw.copy(), cov=data["mm_cov"].data.copy(),
                   binmap=data["mm_binmap"].data.copy(), map=data["mm_map"].data.copy(),
                   amps=amps.local.copy(), n_global=amps.n_global, history=list(mapper.history),
                   g2l=data["pixel_dist"].global_submap_to_local.copy())
        gathered = [None] * world
        dist.all_gather_object(gathered, res)
        if rank == 0:
            out.put(gathered)
    finally:
        mp_.undo()
        dist.destroy_process_group()


def test_two_rank_mapmaker_matches_the_single_rank_oracle():
    world, n_det_total, n_samp = 2, 4, 6000
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_mapmaker_worker, args=(r, world, port, n_det_total, n_samp, out))
             for r in range(world)]
    for p in procs:
        p.start()
    ranks = out.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    full = S.make_observation("c1", n_det=n_det_total, n_samp=n_samp, nside=32, eps_max=0.03)
    pb = O.build_problem(full, O, rcond_threshold=1.0e-3)
    rhs_ref = O.solver_rhs(pb, O, full["signal"])
    amps_ref, hist_ref = O.solve(pb, O, rhs_ref, convergence=1e-30, n_iter_max=4, n_iter_min=4)
    clean = full["signal"].copy()
    O.template_add(pb, O, -amps_ref, clean)
    per = pb.n_amp // world
    for r, got in enumerate(ranks):
        np.testing.assert_array_equal(got["g2l"], pb.global2local)
        assert got["n_global"] == pb.n_amp
        # integer hit map: exact; sums over two ranks re-associate the floating-point ones
        hits_ref = np.zeros(pb.n_local_submap * pb.n_pix_submap, dtype=np.int64)
        sf0 = ((full["det_flags"] & 1) != 0) | ((full["shared_flags"] & 1) != 0)[None, :]
        for d in range(n_det_total):
            for iv in pb.intervals:
                a, b = int(iv["first"]), int(iv["last"])
                sm, lp = O.global_to_local(pb.pixels[d, a:b], pb.n_pix_submap, pb.global2local)
                lp[sf0[d, a:b]] = -1
                O.cov_accum_diag_hits(pb.n_local_submap, pb.n_pix_submap, 3, sm, lp, hits_ref)
        np.testing.assert_array_equal(got["hits"], hits_ref)
        # (measured: 1e-14 .. 1e-18; the bars are the north-star 1e-10 and tighter)
        H.assert_close_norm(got["cov"], pb.cov, rtol=1e-12, what=f"covariance (rank {r})")
        H.assert_close_norm(got["binmap"], O.bin_map(pb, O, full["signal"], O.cov_apply_diag),
                            rtol=1e-12, what=f"binned map (rank {r})")
        np.testing.assert_allclose(got["history"], hist_ref, rtol=1e-10)
        H.assert_close_norm(got["amps"], amps_ref[r * per:(r + 1) * per],
                            what=f"amplitudes of rank {r}")
        H.assert_close_norm(got["map"], O.bin_map(pb, O, clean, O.cov_apply_diag),
                            what=f"destriped map (rank {r})")
    np.testing.assert_array_equal(ranks[0]["map"], ranks[1]["map"])
