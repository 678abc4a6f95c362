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
