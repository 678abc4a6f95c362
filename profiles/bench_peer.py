#!/usr/bin/env python
"""Micro-benchmark of the map reduction alone (torchrun, N ranks): fused P2P kernel vs NCCL
all-reduce (+ cov_apply), on a map of the bench workload's size, ranks synchronised before each
call so that no load imbalance is included."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from toast_b200 import kernels as KC  # noqa: E402
from toast_b200.solver import PeerMap, SymmPeerMap  # noqa: E402
from toast_b200 import lib as L  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    n_loc, nps = 4734, 3072
    n_pix = n_loc * nps
    cov = torch.rand(n_pix * 6, device=dev, dtype=torch.float64)
    z = torch.rand(n_pix * 3, device=dev, dtype=torch.float64)
    pm = PeerMap(n_pix, dev)
    pm.tensor.copy_(z)
    res = {}

    def timeit(fn, name, reps=20):
        for _ in range(3):
            fn()
        times = []
        for _ in range(reps):
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        t = torch.tensor([sorted(times)[len(times) // 2]], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name] = float(t.item())

    timeit(lambda: pm.reduce_cov(cov), "fused_p2p_reduce_cov_ms")
    try:
        sm = SymmPeerMap(n_pix, dev)
        sm.tensor.copy_(z)
        timeit(lambda: sm.reduce_cov(cov), "fused_nvls_reduce_cov_ms")
        L.load().tb_peer_set_multimem(0)
        timeit(lambda: sm.reduce_cov(cov), "fused_p2p_on_symm_ms")
        L.load().tb_peer_set_multimem(1)
    except Exception as exc:  # noqa: BLE001
        res["nvls_error"] = repr(exc)[:300]
    timeit(lambda: dist.all_reduce(z), "nccl_allreduce_ms")
    timeit(lambda: KC.cov_apply_diag(n_loc, nps, 3, cov, z), "cov_apply_ms")
    if rank == 0:
        nbytes = n_pix * 24
        res["map_bytes"] = nbytes
        res["world"] = world
        per_dir = nbytes * (world - 1) / world
        res["fused_gbs_per_direction_per_gpu"] = per_dir / (res["fused_p2p_reduce_cov_ms"] * 1e-3) / 1e9
        if "fused_nvls_reduce_cov_ms" in res:
            res["nvls_gbs_per_direction_per_gpu"] = nbytes * (1 + 1 / world) / (
                res["fused_nvls_reduce_cov_ms"] * 1e-3) / 1e9
        res["nccl_busbw_gbs"] = 2 * per_dir / (res["nccl_allreduce_ms"] * 1e-3) / 1e9
        print(json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
