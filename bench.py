#!/usr/bin/env python
"""bench.py -- detector-samples/s per destriper PCG iteration (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c3|c2|c5] [--regen]
    python bench.py --impl reference ...      # the reference's compiled CPU path, bounded sample

A *step* is one PCG iteration of the Offset-template destriper (SolverLHS.apply + the vector
updates of solve(), ops/mapmaker_solve.py:665-746) over the rank's shard of synthetic data:

    pass 1   F a -> noise-weighted binning          (k_bin_xs:  pixel-sorted crossing list)
    NVLink   map reduction + 3x3 pixel covariance   (N > 1: fused peer kernel, pipelined with
             (N = 1: k_cov_apply)                    the passes when that measures faster)
    pass 2   F a - P m -> N^-1 -> F^T               (k_proj_xs: the same sorted list)
    PCG      d.q, x/r/s update, r.r, s.r, new d     (+ one scalar read-back for convergence)

Algorithmic bytes are the reference layout's 33 B / det-sample per pass; the
passes stream ~4.6 B / det-sample of crossing records instead.

Default workload ("c4"): BASELINE.json configs[3] detector-sharded -- 128 detectors x 12 h @
50 Hz (2.76e8 det-samples) per GPU, nside 2048 NEST IQU, 1 s baselines; at N GPUs the job is
N x 128 detectors with the map all-reduced (weak scaling; N = 8 is the full 1024-detector
2.2e9-sample configuration).  Every iteration streams ~5 GB per GPU, far more than the 126 MB L2,
so no explicit L2 flush is needed between steps.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from toast_b200 import synthetic as S  # noqa: E402

METRIC = "detector-samples/s per destriper PCG iteration"
UNIT = "det-samples/s"

# per-GPU shard of each workload: (n_det per GPU, n_samp)
SHARDS = {
    "c4": dict(n_det=128, n_samp=2160000),   # 1024 det / 8 GPUs
    "c3": dict(n_det=2000, n_samp=360000),   # single-GPU destriper config
    "c2": dict(n_det=1000, n_samp=360000),
    "c5": dict(n_det=1000, n_samp=500000),   # 8000 det / 8 GPUs, high-contention patch
}
BYTES_PER_SAMPLE_PASS = 33      # pixel 8 + weights 24 + solver flag 1
BYTES_PER_SAMPLE_ITER = 66
# dram__bytes_read.sum + dram__bytes_write.sum per launch on the default workload, from the
# committed `ncu --set full` captures (profiles/r1_ncu_passes.txt, profiles/r1_ncu_crossings.txt)
NCU_TRAFFIC = {
    "k_lhs_pair<0>": 8.246e9, "k_lhs_pair<1>": 7.572e9,
    "k_bin_xs": 1.480e9, "k_lhs_x<1>": 3.373e9, "k_proj_xs": 1.536e9,
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(SHARDS))
    ap.add_argument("--regen", action="store_true",
                    help="regenerate pointing inside every pass instead of streaming it")
    ap.add_argument("--scale", type=float, default=1.0,
                    help="shrink the per-GPU shard (debugging only; the JSON line says so)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(np.max(smax)) if smax else None,
            "power_w_max": float(np.max(power)) if power else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's compiled kernels (oracle/_ref) or the C restatement, bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_problem(workload, n_det, n_samp):
    from oracle import toast_oracle as O

    ref = O.load_ref()
    kern = ref if ref is not None else O
    kind = "reference" if ref is not None else "port"
    obs = S.make_observation(workload, n_det=n_det, n_samp=n_samp)
    pb = O.build_problem(obs, kern)
    covapply = kern.cov_apply_diag
    return O, kern, kind, obs, pb, covapply


def cpu_iteration(O, kern, pb, covapply, st):
    """One PCG iteration with the reference kernels (ops/mapmaker_solve.py:665-746)."""
    fl = pb.amp_flags
    q = O.solver_lhs(pb, kern, st["d"], covapply)
    alpha = st["delta"] / O.amp_dot(st["d"], q, fl)
    st["x"] += st["d"] * alpha
    st["r"] -= q * alpha
    sqsum = O.amp_dot(st["r"], st["r"], fl)
    kern.template_offset_apply_diag_precond(pb.offset_var, st["r"], fl, st["s"], False)
    delta_new = O.amp_dot(st["s"], st["r"], fl)
    beta = delta_new / st["delta"]
    st["delta"] = delta_new
    st["d"] *= beta
    st["d"] += st["s"]
    return sqsum


def cpu_state(O, kern, pb, covapply, signal):
    rhs = O.solver_rhs(pb, kern, signal, covapply)
    s = np.zeros_like(rhs)
    kern.template_offset_apply_diag_precond(pb.offset_var, rhs, pb.amp_flags, s, False)
    return dict(x=np.zeros_like(rhs), r=rhs.copy(), d=s.copy(), s=s,
                delta=O.amp_dot(s, rhs, pb.amp_flags))


def cpu_sample_shape(workload):
    # ~1e6-1e7 det-samples: a few seconds per iteration on a handful of host cores
    sh = SHARDS[workload]
    return 8, min(sh["n_samp"], 540000)


def run_cpu(workload, steps, warmup):
    n_det, n_samp = cpu_sample_shape(workload)
    cores = len(os.sched_getaffinity(0))
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    O, kern, kind, obs, pb, covapply = cpu_problem(workload, n_det, n_samp)
    st = cpu_state(O, kern, pb, covapply, obs["signal"])
    for _ in range(warmup):
        cpu_iteration(O, kern, pb, covapply, st)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_iteration(O, kern, pb, covapply, st)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    n_good = int(sum(int(iv["last"] - iv["first"]) for iv in pb.intervals)) * n_det
    return dict(value=n_good / dt, unit=UNIT, cores=cores, kind=kind,
                sample=f"{n_det} detectors x {n_samp} samples of workload {workload} "
                       f"({n_good} det-samples per iteration, stored pointing, "
                       f"OMP_NUM_THREADS={os.environ['OMP_NUM_THREADS']})",
                ms_per_step=dt * 1e3)


def config_dict(args, world, n_det, n_samp, nside, extra=None):
    cfg = {
        "workload": f"{args.workload}: BASELINE.json destriper config, per-GPU shard "
                    f"{n_det} det x {n_samp} samples, nside {nside}, Offset step "
                    f"{S.CONFIGS[args.workload]['step_time']} s, IQU, "
                    + ("pointing regenerated per pass" if args.regen else "stored pointing"),
        "detectors_total": n_det * world,
        "det_samples_total": n_det * n_samp * world,
        "parallelism": f"detector-sharded x{world}, NCCL map all-reduce" if world > 1
                       else "single GPU",
        "l2_policy": "no flush: every iteration streams far more than the 126 MB L2 (C4 shard: "
                     "1.3 GB of crossing records per pass, 0.33 GB map, 0.66 GB covariance, "
                     "0.22 GB of amplitude vectors; 5 GB of DRAM traffic per iteration by ncu)",
    }
    if args.scale != 1.0:
        cfg["scaled_down"] = args.scale
    if extra:
        cfg.update(extra)
    return cfg


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res = run_cpu(args.workload, args.steps, args.warmup)
    sh = SHARDS[args.workload]
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, args.gpus, sh["n_det"], sh["n_samp"],
                              S.CONFIGS[args.workload]["nside"]),
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def build_gpu_problem(args, rank, world, device):
    """Untimed setup, all on the device except the (small) boresight / focalplane generation."""
    import torch

    from toast_b200 import kernels as K
    from toast_b200.solver import DeviceObservation, Destriper

    sh = SHARDS[args.workload]
    n_det = max(2, int(sh["n_det"] * args.scale))
    n_samp = max(1000, int(sh["n_samp"] * args.scale))
    obs = S.make_observation(args.workload, n_det=n_det, n_samp=n_samp, det_first=rank * n_det,
                             with_signal=False, flags=False)
    nside, nest = obs["nside"], obs["nest"]
    n_submap, nps = S.n_submap_for(nside, 16)
    iv = obs["intervals"]
    step = obs["step_length"]

    g = torch.Generator(device=device)
    g.manual_seed(20261017 + 97 * rank)
    # shared flags: 0.5 % in bursts of 20; detector flags: 1 % in bursts of 50
    def bursts(shape, frac, burst):
        n = int(np.prod(shape))
        f = torch.zeros(n + burst, dtype=torch.uint8, device=device)
        starts = torch.randint(0, n, (int(frac * n / burst),), generator=g, device=device)
        for k in range(burst):
            f[starts + k] = 1
        return f[:n].reshape(shape).contiguous()

    shared_flags = bursts((n_samp,), 0.005, 20)
    solver_flags = bursts((n_det, n_samp), 0.01, 50)
    in_view = torch.zeros(n_samp, dtype=torch.bool, device=device)
    for v in iv:
        in_view[int(v["first"]):int(v["last"])] = True
    solver_flags |= (~in_view).to(torch.uint8)[None, :]
    solver_flags |= shared_flags[None, :]

    dobs = DeviceObservation(
        focalplane=obs["focalplane"], boresight=obs["boresight"], intervals=iv,
        det_scale=obs["detweight"], step_length=step, nside=nside, nest=nest,
        n_pix_submap=nps, n_submap=n_submap, global2local=np.zeros(n_submap, dtype=np.int64),
        epsilon=obs["epsilon"], gamma=obs["gamma"], cal=obs["cal"], shared_flags=shared_flags,
        shared_flag_mask=1, solver_flags=solver_flags, solver_flag_mask=1, device=device)
    hits = np.zeros(n_submap, dtype=np.uint8)
    dobs.expand_pointing(hits)
    solver_flags |= (dobs.pixels < 0).to(torch.uint8)
    if world > 1:
        ht = torch.from_numpy(hits).to(device)
        torch.distributed.all_reduce(ht, op=torch.distributed.ReduceOp.MAX)
        hits = ht.cpu().numpy()
    local = np.flatnonzero(hits).astype(np.int64)
    g2l = np.full(n_submap, -1, dtype=np.int64)
    g2l[local] = np.arange(len(local))
    dobs.set_global2local(g2l)
    n_loc = len(local)

    # pixel covariance on the device: accumulate, all-reduce, invert (rcond 1e-3)
    idx = np.arange(n_det, dtype=np.int32)
    invcov = torch.zeros((n_loc, nps, 6), dtype=torch.float64, device=device)
    K.cov_accum(g2l, n_loc, nps, 3, None, invcov, idx, dobs.pixels, idx, dobs.weights, idx,
                solver_flags, obs["detweight"], 1, iv, None, 0)
    if world > 1:
        torch.distributed.all_reduce(invcov)
    rcond = torch.zeros(n_loc * nps, dtype=torch.float64, device=device)
    # reference default: MapMaker.solve_rcond_threshold = 1e-8 (ops/mapmaker.py)
    K.cov_invert(n_loc * nps, 3, invcov, rcond, 1.0e-8)
    # rcond mask -> solver flags (scan the bad-pixel map with the I weight = cal = 1)
    bad = torch.zeros((n_loc, nps, 3), dtype=torch.float64, device=device)
    bad[:, :, 0] = (rcond.reshape(n_loc, nps) == 0).to(torch.float64)
    tmp = torch.zeros((n_det, n_samp), dtype=torch.float64, device=device)
    K.ops_scan_map_float64(g2l, nps, bad, tmp, idx, dobs.pixels, idx, dobs.weights, idx, iv, 1.0,
                           True, False, False)
    solver_flags |= (tmp != 0).to(torch.uint8)
    del bad

    # Offset layout + amplitude variance: n_good per step = F^T (good-sample indicator)
    tmp.fill_(1.0)
    n_good = torch.zeros(dobs.n_amp, dtype=torch.float64, device=device)
    zero_flags = torch.zeros(dobs.n_amp, dtype=torch.uint8, device=device)
    K.template_offset_project_signal_batch(idx, tmp, idx, solver_flags, 1, step,
                                           dobs.amp_offsets, dobs.n_amp_views, n_good, zero_flags,
                                           iv)
    amplen = np.concatenate([
        np.minimum(step, int(v["last"] - v["first"]) - step * np.arange(na))
        for v, na in zip(iv, dobs.n_amp_views)]) if len(iv) else np.zeros(0)
    amplen = torch.from_numpy(np.tile(amplen.astype(np.float64), n_det)).to(device)
    keep = (n_good / amplen) > 0.5
    detw = torch.from_numpy(np.repeat(obs["detweight"], dobs.n_amp_det)).to(device)
    offset_var = torch.where(keep, 1.0 / (detw * torch.clamp(n_good, min=1.0)),
                             torch.zeros_like(n_good))
    amp_flags = (~keep).to(torch.uint8)

    # signal: white noise + per-step random-walk baselines (destriping has work to do)
    tmp.normal_(generator=g)
    tmp *= torch.from_numpy(obs["sigma"]).to(device)[:, None]
    base = torch.cumsum(torch.randn((n_det, dobs.n_amp_det), generator=g, device=device,
                                    dtype=torch.float64), dim=1).reshape(-1).contiguous()
    K.template_offset_add_to_signal_batch(step, dobs.amp_offsets, dobs.n_amp_views, base,
                                          zero_flags, idx, tmp, iv)
    signal = tmp

    ds = Destriper([dobs], n_loc, nps, invcov, offset_var, amp_flags, regen=args.regen,
                   device=device)
    n_good_samples = int(sum(int(v["last"] - v["first"]) for v in iv)) * n_det
    info = dict(n_det=n_det, n_samp=n_samp, nside=nside, n_local_submap=n_loc,
                n_amp=dobs.n_amp, det_samples=n_good_samples,
                flagged_fraction=float((solver_flags != 0).float().mean().item()))
    return ds, dobs, signal, info


def main_gpu(args):
    import torch

    from toast_b200 import lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the toast_b200 path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=device)
    lib = L.load()

    ds, dobs, signal, info = build_gpu_problem(args, rank, world, device)
    from toast_b200.solver import _PCGState

    # PCG state exactly as solve() leaves it before the loop
    n = ds.n_amp
    st = _PCGState(n, device)
    rhs = ds.rhs([signal])
    st.r.copy_(rhs)
    L.check(lib.tb_template_offset_apply_diag_precond(
        L.ptr(ds.offset_var), L.ptr(st.r), L.ptr(ds.amp_flags), L.ptr(st.s), n, L.TB_MEM_DEVICE,
        None))
    st.d.copy_(st.s)
    ds.dot(st.d, st.r, st.delta)
    tmp = torch.zeros(1, dtype=torch.float64, device=device)
    ds.dot(rhs, rhs, tmp)
    sqsum_init = float(tmp.item())

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    ev = lambda: torch.cuda.Event(enable_timing=True)
    history = []

    host_sums = torch.zeros(2, dtype=torch.float64).pin_memory()
    ev_sums = torch.cuda.Event()

    def lhs_and_dot(timers=None):
        # q = A d (pass 1, map reduction + covariance, pass 2) and d.q -- Destriper.lhs, the
        # call solve() makes; with N > 1 it pipelines the passes with the map reduction
        ds.lhs(st.d, st.q, timers)
        ds.dot(st.d, st.q, st.dq)

    def step(timers=None):
        # One PCG iteration exactly as Destriper.solve runs it: x/r/s update + r.r, s.r of the
        # current iteration, then -- while the host reads r.r back for the convergence test --
        # the new direction and the NEXT iteration's LHS (both passes + reduction) and d.q.
        ds.update(st)
        host_sums.copy_(st.sums, non_blocking=True)
        ev_sums.record()
        ds.advance_direction(st)
        lhs_and_dot(timers)
        ev_sums.synchronize()
        history.append(float(host_sums[0]) / sqsum_init)  # host convergence test

    lhs_and_dot()  # the LHS of the first iteration (prologue of the pipelined loop)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # sampled under load: warm-up + timed region + e2e region
    for _ in range(args.warmup):
        step()

    timers = []
    launches0 = lib.tb_launch_count()
    barrier()
    t_start, t_end = ev(), ev()
    t_start.record()
    for _ in range(args.steps):
        step(None if ds.pipeline else timers)
    t_end.record()
    barrier()
    launches = lib.tb_launch_count() - launches0
    if ds.pipeline and getattr(ds, "use_graph", False):
        # the pipelined LHS is replayed from a CUDA graph, which tb_launch_count does not see:
        # per LHS the amplitude prescale + (pass 1, ranged reduction, pass 2) per chunk
        launches += args.steps * (1 + 3 * int(ds.n_chunks))
    ms_total = t_start.elapsed_time(t_end)
    ms_step = ms_total / max(args.steps, 1)
    if ds.pipeline and os.environ.get("TB_PIPE_TIMELINE", "0") == "1":
        # diagnostics: when every launch of one pipelined LHS starts and ends (ms from the first)
        for rep in range(3):
            tl = []
            barrier()
            ds._enqueue_pipelined(st.d, st.q, tl)
            barrier()
        if rank == 0:
            t0 = tl[0][1]
            for label, e0, e1 in tl:
                print(f"timeline {label:12s} {t0.elapsed_time(e0):8.3f} -> "
                      f"{t0.elapsed_time(e1):8.3f} ms", file=sys.stderr)
    if ds.pipeline:
        # per-phase durations come from a few extra UN-pipelined applications of the same LHS
        # (outside the timed region; they change neither x nor r)
        for _ in range(5):
            lhs_and_dot(timers)
        barrier()
    p1 = float(np.mean([e[0].elapsed_time(e[1]) for e in timers]))
    p2 = float(np.mean([e[2].elapsed_time(e[3]) for e in timers]))
    pr = float(np.mean([e[1].elapsed_time(e[2]) for e in timers]))

    # ---- end to end: the LHS through the host-facing call, amplitudes in pinned host memory ----
    d_host = torch.empty(n, dtype=torch.float64).pin_memory()
    q_host = torch.empty(n, dtype=torch.float64).pin_memory()
    d_host.copy_(st.d)
    d_dev = torch.empty(n, dtype=torch.float64, device=device)
    q_dev = torch.empty(n, dtype=torch.float64, device=device)

    def e2e_step():
        d_dev.copy_(d_host, non_blocking=True)
        ds.lhs(d_dev, q_dev)
        q_host.copy_(q_dev, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(min(args.warmup, 2)):
        e2e_step()
    barrier()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1) / max(args.steps, 1)
    clocks = sampler.stop() if rank == 0 else None

    # max over ranks
    if world > 1:
        t = torch.tensor([ms_step, ms_e2e, p1, p2, pr], dtype=torch.float64, device=device)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms_step, ms_e2e, p1, p2, pr = [float(x) for x in t.tolist()]
        cnt = torch.tensor([info["det_samples"]], dtype=torch.float64, device=device)
        torch.distributed.all_reduce(cnt)
        total_samples = float(cnt.item())
    else:
        total_samples = float(info["det_samples"])

    if rank == 0:
        peaks = {}
        pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk_path):
            peaks = json.load(open(pk_path))
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else \
            "fallback 6650 GB/s (B200_PROFILING.md)"
        compact = dobs.has_compact_pointing() and not args.regen
        pair = compact and lib.tb_get_option(b"pair") == 1 and lib.tb_get_option(b"compact") == 1
        pairw = pair and lib.tb_get_option(b"pairw") == 1 and \
            bool(lib.tb_obs_has_pair_weights(dobs.handle().h))
        import ctypes as ct
        n_rec, n_rows, xp = ct.c_int64(0), ct.c_int64(0), ct.c_int(0)
        lib.tb_obs_crossing_stats(dobs.handle().h, ct.byref(n_rec), ct.byref(n_rows), ct.byref(xp))
        crossings = compact and lib.tb_get_option(b"crossings") == 1 and n_rec.value > 0
        blocked = ds._blocked()
        fused = blocked and world == 1 and ds.fuse_lhs
        if blocked:
            names = ("k_bx<0> (pass 1: template -> noise-weighted map, block-ordered crossing "
                     "list, shared-memory map tiles)",
                     "k_bx<2> (pass 1 + covariance + pass 2 fused, block-ordered crossing list, "
                     "shared-memory map tiles)" if fused else
                     "k_bx<1> (pass 2: scan - weight - project, block-ordered crossing list)")
        elif crossings:
            sp = int(lib.tb_obs_sorted_passes(dobs.handle().h))
            names = ("k_bin_xs (pass 1: template -> noise-weighted map, pixel-sorted crossing list)"
                     if sp >= 1 else
                     "k_lhs_x<0> (pass 1: template -> noise-weighted map, crossing list)",
                     "k_proj_xs (pass 2: scan - weight - project, pixel-sorted crossing list)"
                     if sp == 2 else
                     "k_lhs_x<1> (pass 2: scan - weight - project, crossing list)")
        elif pairw:
            names = ("k_lhs_pairw<0> (pass 1: template -> noise-weighted map)",
                     "k_lhs_pairw<1> (pass 2: scan - weight - project)")
        elif pair:
            names = ("k_lhs_pair<0> (pass 1: template -> noise-weighted map)",
                     "k_lhs_pair<1> (pass 2: scan - weight - project)")
        elif compact:
            names = ("k_lhs_compact<0> (pass 1: template -> noise-weighted map)",
                     "k_lhs_compact<1> (pass 2: scan - weight - project)")
        elif args.regen:
            names = ("k_bin<REGEN> (pass 1)", "k_project<REGEN> (pass 2)")
        else:
            names = ("k_bin (pass 1)", "k_project (pass 2)")
        dom, dom_ms = (names[0], p1) if p1 >= p2 else (names[1], p2)
        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed
        # `ncu --set full` capture of THIS workload (profiles/r1_ncu_passes.txt); null otherwise
        traffic = None
        if args.workload == "c4" and args.scale == 1.0 and world == 1:
            traffic = NCU_TRAFFIC.get(dom.split(" ")[0])
        alg_bytes = info["det_samples"] * BYTES_PER_SAMPLE_PASS
        achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
        iter_gbs = info["det_samples"] * BYTES_PER_SAMPLE_ITER / (ms_step * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": total_samples / (ms_step * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": config_dict(args, world, info["n_det"], info["n_samp"], info["nside"],
                                  {"n_amplitudes_per_gpu": info["n_amp"],
                                   "n_local_submaps": info["n_local_submap"],
                                   "flagged_fraction": round(info["flagged_fraction"], 4)}),
            "roofline": {
                "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes,
                "pass1_ms": p1, "pass2_ms": p2, "reduce_cov_ms": pr,
                "map_reduction_tuning_ms": getattr(ds.peer, "tune_ms", None),
                "map_reduction": ("fused NVLS kernel: multimem.ld_reduce + cov + multimem.st"
                                  if getattr(ds.peer, "use_multimem", False) else
                                  "fused P2P reduce-scatter + cov + all-gather kernel")
                                 if ds.peer is not None else
                                 ("NCCL all-reduce + cov_apply" if world > 1 else "cov_apply"),
                "zmap_bytes": int(ds.zmap.numel() * 8),
                "pipeline_tuning_ms": getattr(ds, "pipe_tune_ms", None),
                "pipeline": (f"{ds.n_chunks} pixel chunks: pass 1 -> NVLink reduction -> pass 2 "
                             "overlapped on two streams; pass1/pass2/reduce_cov_ms are from "
                             "un-pipelined applications outside the timed region")
                            if ds.pipeline else None,
                "streamed_bytes_per_sample_per_pass":
                    (round(32.0 * n_rec.value / (info["n_det"] * info["n_samp"]), 2) if crossings
                     else (12 if pairw else 20)) if compact else (1 if args.regen else 33),
                "crossing_records": n_rec.value if crossings else None,
                "iteration_effective_gbs_per_gpu": iter_gbs,
                "iteration_frac_of_peak": iter_gbs / peak,
            },
            "e2e": {"value": total_samples / (ms_e2e * 1e-3), "unit": UNIT,
                    "ms_per_step": ms_e2e, "h2d_bytes_per_step": n * 8,
                    "d2h_bytes_per_step": n * 8,
                    "what": "Destriper.lhs (SolverLHS.apply) with the amplitude vectors in pinned "
                            "host memory: H2D of d, both fused passes + map reduction + "
                            "covariance, D2H of q; timestream data device-resident as in the "
                            "reference's accel pipeline"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "pcg_relative_residuals": history[args.warmup:args.warmup + 5],
        }
        if not args.no_cpu_baseline and world == 1:  # reported at N=1 only
            try:
                res = run_cpu(args.workload, 2, 1)
                line["cpu_baseline"] = {k: res[k] for k in ("value", "unit", "cores", "kind",
                                                            "sample")}
            except Exception as exc:  # the baseline is a report, never a gate
                line["cpu_baseline"] = {"error": str(exc)}
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_gpu(a)
