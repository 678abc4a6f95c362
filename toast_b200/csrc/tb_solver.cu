// tb_solver.cu -- the per-PCG-iteration destriper passes, fused.
//
// One SolverLHS.apply of the reference (ops/mapmaker_solve.py:342-506) runs, per observation,
//   TemplateMatrix.add_to_signal -> [PixelsHealpix, StokesWeights] -> BuildNoiseWeighted
//   ... zmap sync, covariance_apply ...
//   add_to_signal -> [pixels, weights] -> ScanMap(subtract) -> NoiseWeight -> project_signal
// as 7-9 separate kernels around a `det_temp` timestream that is zeroed and refilled twice.
// Here each half is ONE kernel and det_temp lives in a register:
//   pass 1 (k_bin)      zmap[pix] += (F a)[s] * w_det * weights[s]
//   pass 2 (k_project)  out[amp(s)] += w_det * ((F a)[s] - sum_k weights[s,k] * map[pix,k])
// With stored pointing a pass streams 33 B per det-sample (pixel 8 + weights 24 + flag 1); with
// REGEN the pointing is recomputed from the boresight in registers and only the flag byte is read.
#include "tb_device.cuh"
#include "tb_runtime.cuh"

using namespace tbd;

struct tb_obs {
    void *blob = nullptr; // one device allocation holding every small array
    Views V;
    // device arrays inside blob
    const int64_t *amp_view_off = nullptr;
    const int64_t *amp_offsets = nullptr;
    const int64_t *g2l = nullptr;
    const double *fp = nullptr;      // [n_det,4]
    const double *cal = nullptr;
    const double *eta = nullptr;     // (1-eps)/(1+eps)
    const double *gamma = nullptr;
    const double *det_scale = nullptr;
    tb_obs_desc d; // copy of the descriptor (host pointers in it are NOT kept alive)
    int64_t n_amp_det = 0;
};

namespace {

struct ObsDev {
    Views V;
    int64_t n_det, n_samp;
    const int64_t *amp_view_off, *amp_offsets, *g2l;
    const double *fp, *cal, *eta, *gamma, *det_scale;
    double inv_step, inv_nps;
    int64_t n_pix_submap;
    tbm::PixCtx ctx;
    double U_sign;
    const double *boresight;
    const uint8_t *shared_flags;
    uint8_t shared_mask;
    const uint8_t *solver_flags;
    uint8_t solver_mask;
    const int64_t *pixels;
    const double *weights;
    const double *hwp;
};

__device__ unsigned long long g_exact_count_solver = 0ull;

#ifndef TB_BIN_RUN_CAP
#define TB_BIN_RUN_CAP 8
#endif
constexpr int kBinRunCap = TB_BIN_RUN_CAP; // see find_runs<CAP>

// pixel + weights of one sample, either streamed from HBM or regenerated from the boresight
template <bool REGEN, bool NEST>
__device__ __forceinline__ void sample_pointing(const ObsDev &o, int det, int64_t s, bool need,
                                                int64_t &pix, double &w0, double &w1, double &w2,
                                                int &n_exact) {
    if (!REGEN) {
        int64_t i = (int64_t)det * o.n_samp + s;
        pix = ld_stream(o.pixels + i);
        const double *w = o.weights + 3 * i;
        w0 = ld_stream(w);
        w1 = ld_stream(w + 1);
        w2 = ld_stream(w + 2);
    } else {
        pix = -1;
        w0 = w1 = w2 = 0.0;
        bool bad = o.shared_flags ? ((__ldg(o.shared_flags + s) & o.shared_mask) != 0) : false;
        if (bad || !need) return;
        tbm::Quat f = ld_quat(o.fp + 4 * det);
        tbm::Quat q = tbm::qmul(ld_quat(o.boresight + 4 * s), f);
        double dx, dy, dz;
        tbm::rot_zaxis(q, dx, dy, dz);
        int ex = 0;
        pix = tbm::vec2pix<NEST>(o.ctx, dx, dy, dz, &ex);
        n_exact += ex;
        double cal = __ldg(o.cal + det), eta = __ldg(o.eta + det);
        if (o.hwp) {
            tbm::stokes_iqu<true>(q, cal, eta, o.U_sign, __ldg(o.gamma + det), __ldg(o.hwp + s), w0,
                                  w1, w2);
        } else {
            tbm::stokes_iqu<false>(q, cal, eta, o.U_sign, 0.0, 0.0, w0, w1, w2);
        }
    }
}

#define TBS_TILE_LOOP(o)                                                                   \
    TileId _tile = tile_of_block(blockIdx.x, (o).n_det);                                   \
    const int det = _tile.det;                                                             \
    const int lane = threadIdx.x & 31;                                                     \
    _Pragma("unroll") for (int _k = 0; _k < kPerThread; ++_k)

#define TBS_COORDS(o)                                                                      \
    int64_t _t = _tile.t0 + (int64_t)_k * kThreads + threadIdx.x;                          \
    bool valid = _t < (o).V.total;                                                         \
    int view = 0;                                                                          \
    int64_t off = 0, s = 0;                                                                \
    if (valid) {                                                                           \
        view = ((o).V.n_view > 1) ? find_view((o).V, _t) : 0;                              \
        off = _t - __ldg((o).V.prefix + view);                                             \
        s = __ldg((o).V.first + view) + off;                                               \
    }

// ---- pass 1: template -> timestream -> noise-weighted map -------------------------------------
// FROM_SIGNAL: bin a stored timestream instead of the template amplitudes (RHS / final BinMap).
template <bool REGEN, bool NEST, bool FROM_SIGNAL>
__global__ void __launch_bounds__(kThreads)
k_bin(ObsDev o, const double *__restrict__ amps, const uint8_t *__restrict__ aflags,
      const double *__restrict__ signal, double *__restrict__ zmap) {
    int n_exact = 0;
    TBS_TILE_LOOP(o) {
        TBS_COORDS(o)
        int64_t key = -1;
        double z0 = 0.0, z1 = 0.0, z2 = 0.0;
        if (valid) {
            int64_t i = (int64_t)det * o.n_samp + s;
            bool ok = o.solver_flags ? ((ld_stream(o.solver_flags + i) & o.solver_mask) == 0) : true;
            double tod = 0.0;
            if (FROM_SIGNAL) {
                if (ok) tod = ld_stream(signal + i);
            } else {
                int64_t amp = __ldg(o.amp_offsets + det) + __ldg(o.amp_view_off + view) +
                              fast_div(off, o.inv_step);
                if (__ldg(aflags + amp) == 0) tod = __ldg(amps + amp);
            }
            int64_t pix;
            double w0, w1, w2;
            sample_pointing<REGEN, NEST>(o, det, s, ok, pix, w0, w1, w2, n_exact);
            if (ok && pix >= 0) {
                int64_t gsm = fast_div(pix, o.inv_nps);
                key = __ldg(o.g2l + gsm) * o.n_pix_submap + (pix - gsm * o.n_pix_submap);
                double sd = tod * __ldg(o.det_scale + det);
                z0 = sd * w0;
                z1 = sd * w1;
                z2 = sd * w2;
            }
        }
        Runs r = find_runs<kBinRunCap>(key, lane);
        z0 = seg_sum<kBinRunCap>(z0, r);
        z1 = seg_sum<kBinRunCap>(z1, r);
        z2 = seg_sum<kBinRunCap>(z2, r);
        if (r.is_tail && key >= 0) {
            double *z = zmap + key * 3;
            atomicAdd(z, z0);
            atomicAdd(z + 1, z1);
            atomicAdd(z + 2, z2);
        }
    }
    if (REGEN && n_exact) atomicAdd(&g_exact_count_solver, (unsigned long long)n_exact);
}

// ---- pass 2: (template - scanned map) -> noise weight -> template projection ------------------
template <bool REGEN, bool NEST, bool FROM_SIGNAL>
__global__ void __launch_bounds__(kThreads)
k_project(ObsDev o, const double *__restrict__ amps, const uint8_t *__restrict__ aflags,
          const double *__restrict__ signal, const double *__restrict__ binned,
          double *__restrict__ amps_out) {
    int n_exact = 0;
    TBS_TILE_LOOP(o) {
        TBS_COORDS(o)
        int64_t key = -1;
        double v = 0.0;
        if (valid) {
            int64_t i = (int64_t)det * o.n_samp + s;
            int64_t amp = __ldg(o.amp_offsets + det) + __ldg(o.amp_view_off + view) +
                          fast_div(off, o.inv_step);
            bool amp_ok = __ldg(aflags + amp) == 0;
            bool ok = o.solver_flags ? ((ld_stream(o.solver_flags + i) & o.solver_mask) == 0) : true;
            if (amp_ok) key = amp;
            // flagged samples contribute exactly 0 (template_offset.cpp:312-321), so nothing
            // else needs to be read for them
            bool need = amp_ok && ok;
            double tod = 0.0;
            if (FROM_SIGNAL) {
                if (need) tod = ld_stream(signal + i);
            } else {
                if (amp_ok) tod = __ldg(amps + amp);
            }
            int64_t pix;
            double w0, w1, w2;
            sample_pointing<REGEN, NEST>(o, det, s, need, pix, w0, w1, w2, n_exact);
            if (need) {
                if (pix >= 0) {
                    int64_t gsm = fast_div(pix, o.inv_nps);
                    const double *m = binned + 3 * (__ldg(o.g2l + gsm) * o.n_pix_submap +
                                                    (pix - gsm * o.n_pix_submap));
                    double sc = 0.0; // ops_scan_map.cpp:59-64
                    sc += w0 * __ldg(m);
                    sc += w1 * __ldg(m + 1);
                    sc += w2 * __ldg(m + 2);
                    tod -= sc;
                }
                v = tod * __ldg(o.det_scale + det);
            }
        }
        Runs r = find_runs(key, lane);
        v = seg_sum(v, r);
        if (r.is_tail && key >= 0) atomicAdd(amps_out + key, v);
    }
    if (REGEN && n_exact) atomicAdd(&g_exact_count_solver, (unsigned long long)n_exact);
}

ObsDev make_dev(const tb_obs *obs, int regen) {
    const tb_obs_desc &d = obs->d;
    ObsDev o;
    o.V = obs->V;
    o.n_det = d.n_det;
    o.n_samp = d.n_samp;
    o.amp_view_off = obs->amp_view_off;
    o.amp_offsets = obs->amp_offsets;
    o.g2l = obs->g2l;
    o.fp = obs->fp;
    o.cal = obs->cal;
    o.eta = obs->eta;
    o.gamma = obs->gamma;
    o.det_scale = obs->det_scale;
    o.inv_step = 1.0 / (double)d.step_length;
    o.inv_nps = 1.0 / (double)d.n_pix_submap;
    o.n_pix_submap = d.n_pix_submap;
    o.ctx = tbm::make_pix_ctx(d.nside, 1.0);
    o.U_sign = d.IAU ? -1.0 : 1.0;
    o.boresight = d.boresight;
    o.shared_flags = d.shared_flags;
    o.shared_mask = d.shared_flag_mask;
    o.solver_flags = d.solver_flags;
    o.solver_mask = d.solver_flag_mask;
    o.pixels = d.pixels;
    o.weights = d.weights;
    o.hwp = d.hwp;
    if (regen) {
        TB_REQUIRE(d.boresight != nullptr && d.focalplane != nullptr,
                   "regen needs boresight and focalplane");
    } else {
        TB_REQUIRE(d.pixels != nullptr && d.weights != nullptr,
                   "stored-pointing pass needs pixels and weights");
    }
    return o;
}

inline int64_t obs_blocks(const tb_obs *obs) {
    int64_t tiles = (obs->V.total + kTile - 1) / kTile;
    return tiles * obs->d.n_det;
}

#define TBS_LAUNCH(kernel, nb, stream, ...)                                                \
    do {                                                                                   \
        if ((nb) > 0) {                                                                    \
            TB_REQUIRE((nb) < 2147483647LL, "grid too large");                             \
            kernel<<<(unsigned)(nb), kThreads, 0, (cudaStream_t)(stream)>>>(__VA_ARGS__);  \
            TB_CUDA(cudaGetLastError());                                                   \
            tbr::count_launch();                                                           \
        }                                                                                  \
    } while (0)

template <bool FROM_SIGNAL>
void launch_bin(const tb_obs *obs, const double *amps, const uint8_t *aflags, const double *signal,
                double *zmap, int regen, void *stream) {
    ObsDev o = make_dev(obs, regen);
    int64_t nb = obs_blocks(obs);
    if (!regen) {
        auto k = k_bin<false, true, FROM_SIGNAL>;
        TBS_LAUNCH(k, nb, stream, o, amps, aflags, signal, zmap);
    } else if (obs->d.nest) {
        auto k = k_bin<true, true, FROM_SIGNAL>;
        TBS_LAUNCH(k, nb, stream, o, amps, aflags, signal, zmap);
    } else {
        auto k = k_bin<true, false, FROM_SIGNAL>;
        TBS_LAUNCH(k, nb, stream, o, amps, aflags, signal, zmap);
    }
}

template <bool FROM_SIGNAL>
void launch_project(const tb_obs *obs, const double *amps, const uint8_t *aflags,
                    const double *signal, const double *binned, double *out, int regen,
                    void *stream) {
    ObsDev o = make_dev(obs, regen);
    int64_t nb = obs_blocks(obs);
    if (!regen) {
        auto k = k_project<false, true, FROM_SIGNAL>;
        TBS_LAUNCH(k, nb, stream, o, amps, aflags, signal, binned, out);
    } else if (obs->d.nest) {
        auto k = k_project<true, true, FROM_SIGNAL>;
        TBS_LAUNCH(k, nb, stream, o, amps, aflags, signal, binned, out);
    } else {
        auto k = k_project<true, false, FROM_SIGNAL>;
        TBS_LAUNCH(k, nb, stream, o, amps, aflags, signal, binned, out);
    }
}

// ---- amplitude-vector kernels ------------------------------------------------------------------
constexpr int kRedBlocks = 592; // 4 x 148 SMs
struct RedScratch {
    double *partials = nullptr; // [2 * kRedBlocks]
    unsigned int *counter = nullptr;
};
RedScratch g_red[64];

RedScratch &red_scratch() {
    int dev = 0;
    TB_CUDA(cudaGetDevice(&dev));
    RedScratch &r = g_red[dev & 63];
    if (r.partials == nullptr) {
        TB_CUDA(cudaMalloc(&r.partials, sizeof(double) * 2 * kRedBlocks));
        TB_CUDA(cudaMalloc(&r.counter, sizeof(unsigned int)));
        TB_CUDA(cudaMemset(r.counter, 0, sizeof(unsigned int)));
    }
    return r;
}

__device__ __forceinline__ double block_sum(double v, double *sh) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) {
        for (int w = 0; w < kThreads / 32; ++w) t += sh[w];
    }
    return t; // valid in thread 0
}

// Deterministic grid reduction: each CTA writes its partial(s); the last CTA to finish adds them
// in index order.  NRED = 1 or 2 results.
template <int NRED>
__device__ __forceinline__ void grid_finish(double p0, double p1, double *partials,
                                            unsigned int *counter, double *out) {
    __shared__ bool last;
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = p0;
        if (NRED == 2) partials[kRedBlocks + blockIdx.x] = p1;
        __threadfence();
        unsigned int done = atomicAdd(counter, 1u);
        last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (last && threadIdx.x < 32) {
        __threadfence();
        for (int r = 0; r < NRED; ++r) {
            double acc = 0.0;
            // fixed order: lane-strided partial sums, then a fixed shuffle tree
            for (unsigned int b = threadIdx.x; b < gridDim.x; b += 32)
                acc += ((volatile double *)partials)[r * kRedBlocks + b];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
            if (threadIdx.x == 0) out[r] = acc;
        }
        if (threadIdx.x == 0) *counter = 0u;
    }
}

__global__ void __launch_bounds__(kThreads)
k_amp_dot(const double *__restrict__ a, const double *__restrict__ b,
          const uint8_t *__restrict__ flags, int64_t n, double *partials, unsigned int *counter,
          double *out) {
    __shared__ double sh[kThreads / 32];
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * kThreads) {
        if (flags == nullptr || flags[i] == 0) acc += a[i] * b[i];
    }
    double t = block_sum(acc, sh);
    grid_finish<1>(t, 0.0, partials, counter, out);
}

__global__ void __launch_bounds__(kThreads)
k_pcg_update(const double *__restrict__ delta, const double *__restrict__ dq,
             double *__restrict__ x, double *__restrict__ r, const double *__restrict__ d,
             const double *__restrict__ q, double *__restrict__ s,
             const double *__restrict__ var, const uint8_t *__restrict__ flags, int64_t n,
             double *partials, unsigned int *counter, double *sums) {
    __shared__ double sh[kThreads / 32];
    const double alpha = delta[0] / dq[0];
    double rr = 0.0, sr = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * kThreads) {
        // mapmaker_solve.py:681-694: temp = d * alpha; x += temp; temp = q * alpha; r -= temp
        double xi = x[i] + d[i] * alpha;
        double ri = r[i] - q[i] * alpha;
        x[i] = xi;
        r[i] = ri;
        bool good = flags[i] == 0;
        double si = good ? ri * var[i] : 0.0;
        s[i] = si;
        if (good) {
            rr += ri * ri;
            sr += si * ri;
        }
    }
    double t0 = block_sum(rr, sh);
    double t1 = block_sum(sr, sh);
    grid_finish<2>(t0, t1, partials, counter, sums);
}

__global__ void __launch_bounds__(kThreads)
k_pcg_direction(const double *__restrict__ dnew, const double *__restrict__ dold,
                double *__restrict__ d, const double *__restrict__ s, int64_t n) {
    const double beta = dnew[0] / dold[0];
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * kThreads) {
        d[i] = d[i] * beta + s[i]; // proposal *= beta; proposal += precond
    }
}

inline int red_grid(int64_t n) {
    int64_t b = (n + kThreads - 1) / kThreads;
    if (b < 1) b = 1;
    if (b > kRedBlocks) b = kRedBlocks;
    return (int)b;
}

} // namespace

extern "C" {

tb_obs *tb_obs_create(const tb_obs_desc *desc) {
    try {
        tbr::require_device();
        TB_REQUIRE(desc != nullptr, "NULL descriptor");
        const tb_obs_desc &d = *desc;
        TB_REQUIRE(d.n_det > 0 && d.n_samp > 0 && d.n_view >= 0, "bad observation shape");
        TB_REQUIRE(d.step_length > 0, "step_length must be positive");
        TB_REQUIRE(d.nside > 0 && (d.nside & (d.nside - 1)) == 0 && d.nside <= (1 << 24),
                   "nside must be a power of two <= 2^24");
        TB_REQUIRE(d.global2local && d.det_scale && d.amp_offsets && d.n_amp_views && d.intervals,
                   "missing descriptor arrays");
        tb_obs *o = new tb_obs();
        o->d = d;
        // pack: int64 [first(nv), prefix(nv+1), amp_view_off(nv), amp_offsets(nd), g2l(ns)]
        //       double [fp(4nd), cal, eta, gamma, det_scale]
        int64_t nv = d.n_view, nd = d.n_det, ns = d.n_submap;
        std::vector<int64_t> ib(3 * nv + 1 + nd + ns);
        int64_t total = 0, acc = 0;
        for (int64_t v = 0; v < nv; ++v) {
            int64_t a = d.intervals[v].first, b = d.intervals[v].last;
            TB_REQUIRE(a >= 0 && b <= d.n_samp, "interval outside [0, n_samp)");
            ib[v] = a;
            ib[nv + v] = total;
            if (b > a) total += b - a;
            ib[2 * nv + 1 + v] = acc;
            acc += d.n_amp_views[v];
        }
        ib[2 * nv] = total;
        for (int64_t i = 0; i < nd; ++i) ib[3 * nv + 1 + i] = d.amp_offsets[i];
        for (int64_t i = 0; i < ns; ++i) ib[3 * nv + 1 + nd + i] = d.global2local[i];
        std::vector<double> db(8 * nd, 0.0);
        for (int64_t i = 0; i < nd; ++i) {
            if (d.focalplane)
                for (int k = 0; k < 4; ++k) db[4 * i + k] = d.focalplane[4 * i + k];
            double eps = d.epsilon ? d.epsilon[i] : 0.0;
            db[4 * nd + i] = d.cal ? d.cal[i] : 1.0;
            db[5 * nd + i] = (1.0 - eps) / (1.0 + eps);
            db[6 * nd + i] = d.gamma ? d.gamma[i] : 0.0;
            db[7 * nd + i] = d.det_scale[i];
        }
        // keep the double block 32-byte aligned: fp quaternions are read as double2
        while ((ib.size() * sizeof(int64_t)) % 32 != 0) ib.push_back(0);
        size_t ibytes = ib.size() * sizeof(int64_t), dbytes = db.size() * sizeof(double);
        TB_CUDA(cudaMalloc(&o->blob, ibytes + dbytes));
        TB_CUDA(cudaMemcpy(o->blob, ib.data(), ibytes, cudaMemcpyHostToDevice));
        TB_CUDA(cudaMemcpy((char *)o->blob + ibytes, db.data(), dbytes, cudaMemcpyHostToDevice));
        const int64_t *di = (const int64_t *)o->blob;
        const double *dd = (const double *)((char *)o->blob + ibytes);
        o->V.first = di;
        o->V.prefix = di + nv;
        o->V.n_view = (int)nv;
        o->V.total = total;
        o->amp_view_off = di + 2 * nv + 1;
        o->amp_offsets = di + 3 * nv + 1;
        o->g2l = di + 3 * nv + 1 + nd;
        o->fp = dd;
        o->cal = dd + 4 * nd;
        o->eta = dd + 5 * nd;
        o->gamma = dd + 6 * nd;
        o->det_scale = dd + 7 * nd;
        o->n_amp_det = acc;
        // the host arrays of the descriptor are not retained
        o->d.intervals = nullptr;
        o->d.epsilon = o->d.gamma = o->d.cal = o->d.det_scale = nullptr;
        o->d.amp_offsets = o->d.n_amp_views = o->d.global2local = nullptr;
        return o;
    } catch (const tbr::Error &e) {
        tbr::set_error(e.code, e.msg);
        return nullptr;
    }
}

void tb_obs_destroy(tb_obs *obs) {
    if (obs == nullptr) return;
    if (obs->blob) cudaFree(obs->blob);
    delete obs;
}

int tb_lhs_pass1(const tb_obs *obs, const double *amplitudes, const uint8_t *amp_flags,
                 double *zmap, int regen, void *stream) {
    TB_API_BEGIN
    tbr::require_device();
    TB_REQUIRE(obs && amplitudes && amp_flags && zmap, "NULL argument");
    launch_bin<false>(obs, amplitudes, amp_flags, nullptr, zmap, regen, stream);
    TB_API_END
}

int tb_lhs_pass2(const tb_obs *obs, const double *amplitudes, const uint8_t *amp_flags,
                 const double *binned, double *amplitudes_out, int regen, void *stream) {
    TB_API_BEGIN
    tbr::require_device();
    TB_REQUIRE(obs && amplitudes && amp_flags && binned && amplitudes_out, "NULL argument");
    launch_project<false>(obs, amplitudes, amp_flags, nullptr, binned, amplitudes_out, regen,
                          stream);
    TB_API_END
}

int tb_rhs_project(const tb_obs *obs, const double *signal, const uint8_t *amp_flags,
                   const double *binned, double *amplitudes_out, int regen, void *stream) {
    TB_API_BEGIN
    tbr::require_device();
    TB_REQUIRE(obs && signal && amp_flags && binned && amplitudes_out, "NULL argument");
    launch_project<true>(obs, nullptr, amp_flags, signal, binned, amplitudes_out, regen, stream);
    TB_API_END
}

int tb_bin_signal(const tb_obs *obs, const double *signal, double *zmap, int regen,
                  void *stream) {
    TB_API_BEGIN
    tbr::require_device();
    TB_REQUIRE(obs && signal && zmap, "NULL argument");
    launch_bin<true>(obs, nullptr, nullptr, signal, zmap, regen, stream);
    TB_API_END
}

int tb_amp_dot(const double *a, const double *b, const uint8_t *flags, int64_t n, double *out,
               void *stream) {
    TB_API_BEGIN
    tbr::require_device();
    RedScratch &r = red_scratch();
    k_amp_dot<<<red_grid(n), kThreads, 0, (cudaStream_t)stream>>>(a, b, flags, n, r.partials,
                                                                   r.counter, out);
    TB_CUDA(cudaGetLastError());
    tbr::count_launch();
    TB_API_END
}

int tb_pcg_update(const double *delta, const double *dq, double *x, double *r, const double *d,
                  const double *q, double *s, const double *offset_var, const uint8_t *flags,
                  int64_t n, double *sums, void *stream) {
    TB_API_BEGIN
    tbr::require_device();
    RedScratch &rs = red_scratch();
    k_pcg_update<<<red_grid(n), kThreads, 0, (cudaStream_t)stream>>>(
        delta, dq, x, r, d, q, s, offset_var, flags, n, rs.partials, rs.counter, sums);
    TB_CUDA(cudaGetLastError());
    tbr::count_launch();
    TB_API_END
}

int tb_pcg_direction(const double *delta_new, const double *delta_old, double *d, const double *s,
                     int64_t n, void *stream) {
    TB_API_BEGIN
    tbr::require_device();
    k_pcg_direction<<<red_grid(n), kThreads, 0, (cudaStream_t)stream>>>(delta_new, delta_old, d, s,
                                                                         n);
    TB_CUDA(cudaGetLastError());
    tbr::count_launch();
    TB_API_END
}

} // extern "C"
