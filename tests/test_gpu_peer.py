"""GPU tests of the fused map-reduction + covariance kernel (tb_map_reduce_cov).

world = 1 runs on any GPU box (the kernel degenerates to cov_apply over its own buffer); the
2-rank test spawns two processes when the box has two GPUs and compares the NVLink peer-memory
path with NCCL all-reduce + cov_apply."""

import os
import socket

import numpy as np
import pytest
import torch

import helpers as H  # noqa: F401
from helpers import O, S
from toast_b200 import kernels as KC
from toast_b200 import lib as L

pytestmark = pytest.mark.gpu


def test_single_rank_reduce_cov_equals_cov_apply():
    lib = L.load()
    n_loc, nps = 5, 3072
    n_pix = n_loc * nps
    rng = np.random.default_rng(2)
    cov = rng.standard_normal((n_pix, 6))
    z = rng.standard_normal((n_pix, 3))
    h = lib.tb_peer_create(0, 1, n_pix * 24)
    assert h
    try:
        from toast_b200.solver import _CudaView

        t = torch.as_tensor(_CudaView(lib.tb_peer_map_ptr(h), n_pix * 3), device="cuda")
        t.copy_(torch.from_numpy(z.reshape(-1)))
        cov_d = torch.from_numpy(cov).cuda()
        L.check(lib.tb_map_reduce_cov(h, n_pix, L.ptr(cov_d), None))
        L.check(lib.tb_map_reduce_cov(h, n_pix, L.ptr(cov_d), None))  # twice: epochs advance
        got = t.cpu().numpy().reshape(n_pix, 3)
    finally:
        lib.tb_peer_destroy(h)
    ref = z.reshape(-1).copy()
    O.cov_apply_diag(n_loc, nps, 3, cov.reshape(-1), ref)
    O.cov_apply_diag(n_loc, nps, 3, cov.reshape(-1), ref)
    np.testing.assert_array_equal(got, ref.reshape(n_pix, 3))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        from toast_b200.solver import PeerMap

        n_loc, nps = 67, 3072   # 205 824 pixels: several tiles per rank slice at every world size
        n_pix = n_loc * nps
        g = torch.Generator(device="cuda")
        g.manual_seed(100 + rank)
        z = torch.randn(n_pix * 3, generator=g, device="cuda", dtype=torch.float64)
        g.manual_seed(5)
        cov = torch.randn(n_pix * 6, generator=g, device="cuda", dtype=torch.float64)
        # NCCL path
        ref = z.clone()
        dist.all_reduce(ref)
        KC.cov_apply_diag(n_loc, nps, 3, cov, ref)
        # fused peer path, three rounds to exercise the epoch barriers
        pm = PeerMap(n_pix, torch.device("cuda", rank))
        err = 0.0
        for _ in range(3):
            pm.tensor.copy_(z)
            pm.reduce_cov(cov)
            torch.cuda.synchronize()
            err = max(err, float((pm.tensor - ref).abs().max() / ref.abs().max()))
        # NVLS form on symmetric memory (in-switch reduction + multicast store), and the P2P
        # kernel on the same symmetric buffers
        err_mc = err_sp = -1.0
        try:
            from toast_b200.solver import SymmPeerMap

            sm = SymmPeerMap(n_pix, torch.device("cuda", rank))
        except Exception as exc:  # noqa: BLE001  (no multicast support on this box)
            print("SymmPeerMap unavailable:", repr(exc)[:300], flush=True)
            sm = None
        if sm is not None:
            lib = L.load()
            assert lib.tb_peer_has_multicast(sm.h) == 1
            for mode in (1, 0):
                lib.tb_peer_set_multimem(mode)
                e = 0.0
                for _ in range(3):
                    sm.tensor.copy_(z)
                    sm.reduce_cov(cov)
                    torch.cuda.synchronize()
                    e = max(e, float((sm.tensor - ref).abs().max() / ref.abs().max()))
                if mode == 1:
                    err_mc = e
                else:
                    err_sp = e
            lib.tb_peer_set_multimem(1)
        dist.barrier()
        if rank == 0:
            out.put((err, err_mc, err_sp))
    except BaseException:
        # a rank that fails leaves its peers in a collective or in the device-side barrier, and
        # tearing the process group down can then block for ever: report and leave at once (the
        # parent sees the exit code and stops the other ranks, _collect)
        import sys
        import traceback

        traceback.print_exc()
        sys.stderr.flush()
        os._exit(1)
    finally:
        dist.destroy_process_group()


def _need(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")


def _collect(procs, out, timeout):
    """Result of rank 0, failing FAST: a rank that dies leaves the others in a collective or in
    the device-side barrier of the reduction kernel, so the first non-zero exit code (or the
    deadline) terminates every process instead of waiting for a long queue time-out."""
    import queue
    import time

    deadline = time.monotonic() + timeout
    result = None
    try:
        while result is None:
            try:
                result = out.get(timeout=2.0)
            except queue.Empty:
                bad = [p.exitcode for p in procs if p.exitcode not in (None, 0)]
                assert not bad, f"worker process(es) failed with exit codes {bad}"
                assert time.monotonic() < deadline, f"no result after {timeout} s"
        for p in procs:
            p.join(timeout=60)
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        return result
    finally:
        for p in procs:
            if p.is_alive():
                p.terminate()
        for p in procs:
            p.join(timeout=10)
            if p.is_alive():
                p.kill()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_rank_peer_reduction_matches_nccl(world):
    """tb_map_reduce_cov in its three forms (CUDA-IPC P2P, NVLS multimem on symmetric memory,
    P2P on symmetric memory) against NCCL all-reduce + cov_apply, at every world size the
    box offers (the reference: PixelData.sync_allreduce + covariance_apply, pixels.py:710-779,
    covariance.py:262-306)."""
    import torch.multiprocessing as mp

    _need(world)
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    err, err_mc, err_sp = _collect(procs, out, 120 + 15 * world)
    assert err < 1e-14
    # -1 = the box has no NVLS multicast (reported, not a failure of the kernels)
    assert err_mc < 1e-14 and err_sp < 1e-14


def _solve_shape(world):
    # one detector pair per rank at least; world 2 keeps the round-1 case (3 detectors per rank)
    return (6 if world == 2 else 2 * world), 24000, 64


def _solve_worker(rank, world, port, out):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "1"
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        from toast_b200 import lib as L_
        from toast_b200.solver import DeviceObservation, Destriper

        n_det, n_samp, nside = _solve_shape(world)
        obs = S.make_observation("c4", n_det=n_det, n_samp=n_samp, nside=nside, eps_max=0.03)
        # the full problem (setup stages from the oracle) ...
        pb = O.build_problem(obs, O)
        # ... and this rank's detector shard of it
        per = n_det // world
        sl = slice(rank * per, (rank + 1) * per)
        asl = slice(rank * per * int(pb.n_amp_views.sum()), (rank + 1) * per * int(pb.n_amp_views.sum()))
        dobs = DeviceObservation(
            focalplane=obs["focalplane"][sl], boresight=obs["boresight"],
            intervals=obs["intervals"], det_scale=pb.det_scale[sl], step_length=pb.step_length,
            nside=pb.nside, nest=pb.nest, n_pix_submap=pb.n_pix_submap, n_submap=pb.n_submap,
            global2local=pb.global2local, epsilon=obs["epsilon"][sl], gamma=obs["gamma"][sl],
            cal=obs["cal"][sl], shared_flags=pb.shared_flags, shared_flag_mask=1,
            solver_flags=np.ascontiguousarray(pb.solver_flags[sl]), solver_flag_mask=1,
            device=torch.device("cuda", rank))
        dobs.expand_pointing(np.zeros(pb.n_submap, dtype=np.uint8))
        res = {}
        for fused in (True, False):
            ds = Destriper([dobs], pb.n_local_submap, pb.n_pix_submap, pb.cov,
                           pb.offset_var[asl], pb.amp_flags[asl], fused_reduce=fused,
                           device=torch.device("cuda", rank))
            assert (ds.peer is not None) == fused
            rhs = ds.rhs([torch.from_numpy(np.ascontiguousarray(obs["signal"][sl])).cuda()])
            amps, hist = ds.solve(rhs, n_iter_max=8)
            res[fused] = (rhs.cpu().numpy(), hist)
            if fused:
                # the (opt-in) chunk pipeline -- passes overlapped with the ranged peer
                # reduction on a second stream, replayed from a CUDA graph -- against the same
                # LHS run phase by phase
                ds._setup_pipeline(4)
                from toast_b200.solver import _all_ranks_ok

                assert _all_ranks_ok(ds.pipeline and ds.n_chunks >= 2, ds.device)
                g = torch.Generator(device="cuda")
                g.manual_seed(11 + rank)
                a = torch.randn(ds.n_amp, generator=g, device="cuda", dtype=torch.float64)
                a[ds.amp_flags != 0] = 0.0
                q_pipe, q_ser = torch.zeros_like(a), torch.zeros_like(a)
                ds.prepare_lhs(a, q_pipe)   # every rank captures before any rank replays
                for _ in range(3):
                    ds.lhs(a, q_pipe)
                ds.pipeline = False
                ds.lhs(a, q_ser)
                torch.cuda.synchronize()
                res["pipe_err"] = max(res.get("pipe_err", 0.0),
                                      float((q_pipe - q_ser).abs().max() / q_ser.abs().max()))
        if rank == 0:
            rhs_ref = O.solver_rhs(pb, O, obs["signal"])
            _, hist_ref = O.solve(pb, O, rhs_ref, n_iter_max=8)
            out.put((res[True][0], res[False][0], rhs_ref[asl], res[True][1], res[False][1],
                     hist_ref, res["pipe_err"]))
        dist.barrier()
    except BaseException:
        # a rank that fails leaves its peers in a collective or in the device-side barrier, and
        # tearing the process group down can then block for ever: report and leave at once (the
        # parent sees the exit code and stops the other ranks, _collect)
        import sys
        import traceback

        traceback.print_exc()
        sys.stderr.flush()
        os._exit(1)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_rank_destriper_matches_single_rank_oracle(world):
    """Detector-sharded solve on 2 / 4 / 8 GPUs (fused peer reduction and NCCL) against the
    oracle's single-process solve of the whole problem: RHS to 1e-10, residual history per the
    reproducibility envelope, pipelined LHS == phase-by-phase LHS."""
    import torch.multiprocessing as mp

    _need(world)
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_solve_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    rhs_f, rhs_n, rhs_ref, hist_f, hist_n, hist_ref, pipe_err = _collect(procs, out,
                                                                         150 + 15 * world)
    assert pipe_err < 1e-12, f"pipelined vs phase-by-phase LHS: {pipe_err}"
    H.assert_close_norm(rhs_f, rhs_ref, what="RHS shard (fused)")
    H.assert_close_norm(rhs_n, rhs_ref, what="RHS shard (NCCL)")
    n_det, n_samp, nside = _solve_shape(world)
    obs = S.make_observation("c4", n_det=n_det, n_samp=n_samp, nside=nside, eps_max=0.03)
    pb = O.build_problem(obs, O)
    env = H.pcg_envelope(pb, O.solver_rhs(pb, O, obs["signal"]), 8)
    H.assert_history_matches(hist_f, hist_ref, env, what=f"{world}-rank fused")
    H.assert_history_matches(hist_n, hist_ref, env, what=f"{world}-rank NCCL")
