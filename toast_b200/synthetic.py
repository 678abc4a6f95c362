"""Synthetic focalplanes, scans and timestreams for the five BASELINE.json workloads.

Pure numpy, CPU only, deterministic (``default_rng(20261017 + config_id)``); this is the
input generator of SURVEY.md section 8(d), shared by the parity tests and ``bench.py``.
It replaces the reference's instrument / schedule / simulation layers (out of scope) with
closed-form equivalents:

* quaternion convention ``[x, y, z, w]`` scalar-last (``qarray.py:272-299``),
* satellite boresight: the closed form of ``ops/sim_satellite.py:120-176``,
* ground boresight: ``Rz(RA) Ry(pi/2 - Dec) Rz(pa)`` (``qarray.py:454-485``).
"""

import numpy as np

interval_dtype = np.dtype(
    {
        "names": ["start", "stop", "first", "last"],
        "formats": ["d", "d", "q", "q"],
        "offsets": [0, 8, 16, 24],
    }
)

XAXIS = np.array([1.0, 0.0, 0.0])
YAXIS = np.array([0.0, 1.0, 0.0])
ZAXIS = np.array([0.0, 0.0, 1.0])


def q_rotation(axis, angle):
    """Quaternion(s) for a rotation of ``angle`` about ``axis``."""
    angle = np.asarray(angle, dtype=np.float64)
    half = 0.5 * angle
    s = np.sin(half)
    out = np.empty(angle.shape + (4,), dtype=np.float64)
    out[..., 0] = axis[0] * s
    out[..., 1] = axis[1] * s
    out[..., 2] = axis[2] * s
    out[..., 3] = np.cos(half)
    return out


def q_mult(p, q):
    """Hamilton product p (x) q, broadcasting over leading axes."""
    p = np.asarray(p, dtype=np.float64)
    q = np.asarray(q, dtype=np.float64)
    px, py, pz, pw = p[..., 0], p[..., 1], p[..., 2], p[..., 3]
    qx, qy, qz, qw = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    out = np.empty(np.broadcast(px, qx).shape + (4,), dtype=np.float64)
    out[..., 0] = px * qw + py * qz - pz * qy + pw * qx
    out[..., 1] = -px * qz + py * qw + pz * qx + pw * qy
    out[..., 2] = px * qy - py * qx + pz * qw + pw * qz
    out[..., 3] = -px * qx - py * qy - pz * qz + pw * qw
    return out


def q_norm(q):
    return q / np.sqrt(np.sum(q * q, axis=-1, keepdims=True))


def make_intervals(ranges):
    iv = np.zeros(len(ranges), dtype=interval_dtype)
    for i, (a, b) in enumerate(ranges):
        iv[i] = (float(a), float(b), int(a), int(b))
    return iv


def focalplane(n_det, rng, fov_deg=10.0, eps_max=0.0):
    """Pairs of orthogonal detectors on random pixel positions inside the field of view."""
    n_pix = (n_det + 1) // 2
    theta = np.radians(fov_deg / 2) * np.sqrt(rng.random(n_pix))
    phi = 2 * np.pi * rng.random(n_pix)
    rng.random(n_pix)  # (keeps the stream layout stable)
    psi_pix = np.where(np.arange(n_pix) % 2 == 0, 0.0, np.pi / 4)
    quats = np.zeros((n_det, 4))
    for d in range(n_det):
        p = d // 2
        psi = psi_pix[p] + (np.pi / 2 if (d % 2) else 0.0)
        q = q_mult(q_rotation(ZAXIS, phi[p]),
                   q_mult(q_rotation(YAXIS, theta[p]), q_rotation(ZAXIS, psi - phi[p])))
        quats[d] = q_norm(q)
    epsilon = eps_max * rng.random(n_det)
    gamma = np.zeros(n_det)
    cal = np.ones(n_det)
    sigma = 1.0 + 0.1 * rng.random(n_det)
    detweight = 1.0 / sigma**2
    return quats, epsilon, gamma, cal, sigma, detweight


def satellite_boresight(n_samp, rate, spin_period_s=600.0, spin_angle_deg=30.0,
                        prec_period_s=3000.0, prec_angle_deg=65.0, first=0):
    s = np.arange(first, first + n_samp, dtype=np.float64)
    ph_prec = 2 * np.pi * np.modf(s / (rate * prec_period_s))[0]
    ph_spin = 2 * np.pi * np.modf(s / (rate * spin_period_s))[0]
    q = q_rotation(ZAXIS, np.pi / 2)[None, :]
    q = q_mult(q_rotation(XAXIS, np.radians(spin_angle_deg))[None, :], q)
    q = q_mult(q_rotation(ZAXIS, ph_spin), q)
    q = q_mult(q_rotation(XAXIS, np.radians(prec_angle_deg))[None, :], q)
    q = q_mult(q_rotation(ZAXIS, ph_prec), q)
    q = q_mult(q_rotation(YAXIS, np.pi / 2)[None, :], q)
    return np.ascontiguousarray(q_norm(q))


def ground_boresight(n_samp, rate, ra0_deg=40.0, dec0_deg=-30.0, throw_deg=20.0,
                     drift_deg=10.0, scan_period_s=90.0, pa_deg=15.0):
    t = np.arange(n_samp, dtype=np.float64) / rate
    T = n_samp / rate
    ph = np.modf(t / scan_period_s)[0]
    tri = 2.0 * np.abs(2.0 * ph - 1.0) - 1.0  # [-1, 1] triangle wave
    dec = np.radians(dec0_deg) + np.radians(drift_deg) * (t / T - 0.5)
    ra = np.radians(ra0_deg) + 0.5 * np.radians(throw_deg) * tri / np.cos(dec)
    q = q_mult(q_rotation(ZAXIS, ra),
               q_mult(q_rotation(YAXIS, np.pi / 2 - dec),
                      q_rotation(ZAXIS, np.radians(pa_deg))[None, :]))
    return np.ascontiguousarray(q_norm(q)), ph


def scan_intervals(n_samp, rate, scan_period_s, keep=0.9):
    """Two sweeps per period; the samples around each turnaround are left outside every
    interval so that ~``keep`` of the samples are covered (exercises n_view > 1)."""
    half = scan_period_s * rate / 2.0
    n_half = int(np.ceil(n_samp / half))
    cut = 0.5 * (1.0 - keep) * half
    ranges = []
    for i in range(n_half):
        a = int(np.ceil(i * half + cut))
        b = int(np.floor((i + 1) * half - cut))
        b = min(b, n_samp)
        if b > a:
            ranges.append((a, b))
    return make_intervals(ranges)


def burst_flags(shape, frac, burst, rng, value=1):
    """uint8 flags with ~frac of the samples set, in bursts of ``burst`` samples."""
    flags = np.zeros(shape, dtype=np.uint8)
    flat = flags.reshape(-1)
    n = flat.size
    n_burst = int(frac * n / burst)
    starts = rng.integers(0, max(1, n - burst), size=n_burst)
    for k in range(burst):
        flat[np.minimum(starts + k, n - 1)] = value
    return flags


# name: (config_id, n_det, n_samp, rate, nside, nest, scan, step_time_s)
CONFIGS = {
    "c1": dict(id=1, n_det=4, n_samp=6000, rate=10.0, nside=64, nest=True, scan="satellite",
               step_time=10.0),
    "c2": dict(id=2, n_det=1000, n_samp=360000, rate=100.0, nside=512, nest=True,
               scan="ground", step_time=1.0),
    "c3": dict(id=3, n_det=2000, n_samp=360000, rate=100.0, nside=1024, nest=True,
               scan="ground", step_time=1.0),
    "c4": dict(id=4, n_det=1024, n_samp=2160000, rate=50.0, nside=2048, nest=True,
               scan="satellite", step_time=1.0),
    "c5": dict(id=5, n_det=8000, n_samp=500000, rate=100.0, nside=256, nest=False,
               scan="ground_small", step_time=1.0),
}


def make_observation(name, n_det=None, n_samp=None, det_first=0, with_signal=True,
                     eps_max=0.0, flags=True, nside=None, nest=None):
    """Build one synthetic observation of workload ``name`` ("c1".."c5").

    ``n_det`` / ``n_samp`` shrink the workload (tests); ``det_first`` selects a detector
    block out of the full focalplane (multi-GPU sharding draws the *same* focalplane on
    every rank and slices it).  Returns a dict of numpy arrays in the reference's layouts.
    """
    cfg = dict(CONFIGS[name])
    full_det = cfg["n_det"]
    if n_det is None:
        n_det = full_det
    if n_samp is None:
        n_samp = cfg["n_samp"]
    if nside is not None:
        cfg["nside"] = nside
    if nest is not None:
        cfg["nest"] = nest
    rate = cfg["rate"]
    rng = np.random.default_rng(20261017 + cfg["id"])
    n_fp = max(full_det, det_first + n_det)
    fp, epsilon, gamma, cal, sigma, detweight = focalplane(n_fp, rng, eps_max=eps_max)
    sl = slice(det_first, det_first + n_det)
    fp, epsilon, gamma, cal = fp[sl].copy(), epsilon[sl].copy(), gamma[sl].copy(), cal[sl].copy()
    sigma, detweight = sigma[sl].copy(), detweight[sl].copy()

    if cfg["scan"] == "satellite":
        bore = satellite_boresight(n_samp, rate)
        intervals = make_intervals([(0, n_samp)])
    else:
        small = cfg["scan"] == "ground_small"
        bore, _ = ground_boresight(
            n_samp, rate,
            throw_deg=5.0 if small else 20.0,
            drift_deg=5.0 if small else 10.0,
        )
        intervals = scan_intervals(n_samp, rate, 90.0, keep=0.9)

    out = dict(
        name=name, n_det=n_det, n_samp=n_samp, rate=rate, nside=cfg["nside"],
        nest=cfg["nest"], nside_submap=16, focalplane=np.ascontiguousarray(fp),
        epsilon=epsilon, gamma=gamma, cal=cal, sigma=sigma, detweight=detweight,
        boresight=bore, intervals=intervals, step_time=cfg["step_time"],
        step_length=int(np.rint(cfg["step_time"] * rate)),
    )
    # per-rank RNG stream for the timestream content so detector blocks differ
    rng_d = np.random.default_rng([20261017 + cfg["id"], det_first])
    if flags:
        out["shared_flags"] = burst_flags((n_samp,), 0.005, 20, rng_d)
        out["det_flags"] = burst_flags((n_det, n_samp), 0.01, 50, rng_d)
    else:
        out["shared_flags"] = np.zeros(n_samp, dtype=np.uint8)
        out["det_flags"] = np.zeros((n_det, n_samp), dtype=np.uint8)
    if with_signal:
        step = out["step_length"]
        n_step = (n_samp + step - 1) // step
        sig = rng_d.standard_normal((n_det, n_samp)) * sigma[:, None]
        base = np.cumsum(rng_d.standard_normal((n_det, n_step)), axis=1)
        sig += np.repeat(base, step, axis=1)[:, :n_samp]
        out["signal"] = np.ascontiguousarray(sig)
    return out


def n_submap_for(nside, nside_submap=16):
    """pixels_healpix.py:133-137."""
    nside_submap = min(nside_submap, nside)
    n_pix = 12 * nside * nside
    n_pix_submap = 12 * nside_submap * nside_submap
    n_submap = (nside // nside_submap) ** 2
    assert n_pix_submap * n_submap == n_pix
    return n_submap, n_pix_submap


def sky_value(pix, nside, k):
    """Deterministic pseudo-sky used to put a scanned signal in the timestreams:
    a smooth function of the pixel number (no healpy needed), component k in I,Q,U."""
    x = pix.astype(np.float64) / (12.0 * nside * nside)
    amp = (1.0, 0.1, -0.05)[k]
    return amp * np.cos(37.0 * x + k) * np.sin(11.0 * x * x + 0.5 * k)
