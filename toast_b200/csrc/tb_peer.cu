// tb_peer.cu -- the map reduction of the destriper fused with the pixel-covariance product,
// over NVLink peer memory.
//
// Reference flow per PCG iteration (ops/mapmaker_utils/mapmaker_utils.py:885-925,
// pixels.py:710-779, covariance.py:262-306): copy zmap to the host, MPI all-reduce it, copy it
// back, then run cov_apply over EVERY pixel on EVERY process.  Here one kernel per GPU does
//     reduce-scatter  (sum my slice of the map over all peers with P2P loads)
//  -> covariance apply on that slice only (1/N of the pixels)
//  -> all-gather      (P2P stores of the finished slice into every peer's map)
// with device-side flag barriers before and after, so nothing else is launched and the
// reduction never leaves the NVLink domain.  Map buffers live in cudaMalloc memory shared
// between the ranks' processes through CUDA IPC handles.
#include "tb_device.cuh"
#include "tb_runtime.cuh"

constexpr int kMaxPeers = 16;
constexpr int kPeerTile = 256;             // pixels per tile (768 doubles = 6 KB of smem)
constexpr int kPeerThreads = 128;
constexpr int kPeerPer = 3;                // double2 per thread per tile (128 x 3 x 2 = 768)

struct tb_peer {
    int rank = 0, world = 1;
    size_t map_bytes = 0;
    double *map = nullptr;                 // my map buffer (peer-visible)
    unsigned long long *flags = nullptr;   // [2][kMaxPeers] arrival / done flags (peer-visible)
    double *peer_map[kMaxPeers] = {};
    unsigned long long *peer_flags[kMaxPeers] = {};
    unsigned int *counter = nullptr;       // local block counter
    // barrier epoch, kept on the DEVICE and advanced by the kernel itself so that a captured
    // CUDA graph of the reduction can be replayed
    unsigned long long *epoch_ctr = nullptr;
    bool opened = false;
    // attached form (tb_peer_attach): buffers are owned by the caller (symmetric memory set up by
    // the host plumbing); mc_map is the NVLS multicast address of the same map buffer, or NULL
    bool owned = true;
    double *mc_map = nullptr;
};

namespace {

struct PeerArgs {
    double *map[kMaxPeers];
    unsigned long long *flags[kMaxPeers];
    int rank, world;
    unsigned long long *epoch_ctr;
    unsigned int *counter;
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// One CTA per tile of 256 pixels of MY slice; several CTAs per SM so that the load, compute and
// store phases of different tiles overlap.  WORLD > 0 unrolls the peer loop so that the loads
// from ALL peers are in flight together (peer-load latency is ~2-4 us).
template <int WORLD>
__global__ void __launch_bounds__(kPeerThreads)
k_map_reduce_cov(PeerArgs a, int64_t tile_first, int64_t n_tiles, const double *__restrict__ cov) {
    __shared__ double2 sh[kPeerTile * 3 / 2];
    __shared__ bool is_last;
    const int tid = threadIdx.x;
    const int world = WORLD > 0 ? WORLD : a.world;
    // launches on one peer object are stream-ordered and the counter only moves when the last
    // CTA of a launch retires, so every thread of this launch reads the same value
    const unsigned long long epoch = *reinterpret_cast<volatile unsigned long long *>(a.epoch_ctr) + 1;

    // ---- start barrier: every rank has finished writing its local map (pass 1) ------------
    if (blockIdx.x == 0 && tid < world) {
        __threadfence_system();
        st_release_sys(a.flags[tid] + a.rank, epoch); // row 0: "my map is ready"
    }
    if (tid < world) {
        const unsigned long long *f = a.flags[a.rank] + tid;
        while (ld_acquire_sys(f) < epoch) {
        }
    }
    __syncthreads();

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t e0 = (tile_first + tile) * (int64_t)(kPeerTile * 3 / 2); // double2 index
        double2 acc[kPeerPer];
#pragma unroll
        for (int k = 0; k < kPeerPer; ++k) acc[k] = make_double2(0.0, 0.0);
        if (WORLD > 0) {
            double2 v[WORLD > 0 ? WORLD : 1][kPeerPer];
#pragma unroll
            for (int q = 0; q < WORLD; ++q) {
                const double2 *src = reinterpret_cast<const double2 *>(a.map[q]) + e0;
#pragma unroll
                for (int k = 0; k < kPeerPer; ++k) v[q][k] = __ldcv(src + k * kPeerThreads + tid);
            }
#pragma unroll
            for (int q = 0; q < WORLD; ++q) {
#pragma unroll
                for (int k = 0; k < kPeerPer; ++k) {
                    acc[k].x += v[q][k].x;
                    acc[k].y += v[q][k].y;
                }
            }
        } else {
            for (int q = 0; q < world; ++q) {
                const double2 *src = reinterpret_cast<const double2 *>(a.map[q]) + e0;
#pragma unroll
                for (int k = 0; k < kPeerPer; ++k) {
                    double2 u = __ldcv(src + k * kPeerThreads + tid);
                    acc[k].x += u.x;
                    acc[k].y += u.y;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < kPeerPer; ++k) sh[k * kPeerThreads + tid] = acc[k];
        __syncthreads();
        // covariance apply: thread t owns pixels 2t, 2t+1 of the tile
        {
            double *z = reinterpret_cast<double *>(sh) + 6 * tid;
            const double *c = cov + ((tile_first + tile) * (int64_t)kPeerTile + 2 * tid) * 6;
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const double *m = c + 6 * p;
                double v0 = z[3 * p], v1 = z[3 * p + 1], v2 = z[3 * p + 2];
                double t0 = 0.0, t1 = 0.0, t2 = 0.0; // order of toast_map_cov.cpp:509-517
                t0 += __ldg(m) * v0;
                t0 += __ldg(m + 1) * v1;
                t1 += __ldg(m + 1) * v0;
                t0 += __ldg(m + 2) * v2;
                t2 += __ldg(m + 2) * v0;
                t1 += __ldg(m + 3) * v1;
                t1 += __ldg(m + 4) * v2;
                t2 += __ldg(m + 4) * v1;
                t2 += __ldg(m + 5) * v2;
                z[3 * p] = t0;
                z[3 * p + 1] = t1;
                z[3 * p + 2] = t2;
            }
        }
        __syncthreads();
        // all-gather: coalesced 16-byte stores of the finished tile into every peer's map
        double2 r[kPeerPer];
#pragma unroll
        for (int k = 0; k < kPeerPer; ++k) r[k] = sh[k * kPeerThreads + tid];
#pragma unroll
        for (int q = 0; q < (WORLD > 0 ? WORLD : kMaxPeers); ++q) {
            if (q < world) {
                double2 *dst = reinterpret_cast<double2 *>(a.map[q]) + e0;
#pragma unroll
                for (int k = 0; k < kPeerPer; ++k) dst[k * kPeerThreads + tid] = r[k];
            }
        }
        __syncthreads();
    }

    // ---- end barrier: every rank's slice has landed in my map ------------------------------
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
        unsigned int done = atomicAdd(a.counter, 1u);
        is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        if (tid < world) {
            __threadfence_system();
            st_release_sys(a.flags[tid] + kMaxPeers + a.rank, epoch); // row 1: "I am done"
            const unsigned long long *f = a.flags[a.rank] + kMaxPeers + tid;
            while (ld_acquire_sys(f) < epoch) {
            }
        }
        if (tid == 0) {
            *a.counter = 0u;
            *a.epoch_ctr = epoch;
        }
    }
}

// ---- NVLS variant: the NVSwitch does the reduction and the broadcast ----------------------------
// `mc` is the multicast address of the map (one virtual address bound to the same buffer on every
// GPU of the node).  multimem.ld_reduce pulls one element from ALL GPUs and returns the sum formed
// inside the switch; multimem.st pushes one copy that the switch replicates to ALL GPUs.  Per GPU
// the NVLink traffic falls from 2 (N-1)/N |map| per direction (P2P reduce-scatter + all-gather) to
// about (1 + 1/N) |map|, and the SM issues 1/N of the loads and stores.  fp64 multimem.ld_reduce
// exists as a scalar only (ptxas: no .v2.f64), so loads are 8-byte, perfectly coalesced; the
// stores go out as 16-byte .v4.f32 bit patterns.
__device__ __forceinline__ double mc_ld_reduce_f64(const double *p) {
    double v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st_16(double2 *p, double2 v) {
    unsigned a = (unsigned)__double2loint(v.x), b = (unsigned)__double2hiint(v.x);
    unsigned c = (unsigned)__double2loint(v.y), d = (unsigned)__double2hiint(v.y);
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p),
                 "f"(__uint_as_float(a)), "f"(__uint_as_float(b)), "f"(__uint_as_float(c)),
                 "f"(__uint_as_float(d))
                 : "memory");
}

constexpr int kMcPer = kPeerTile * 3 / kPeerThreads; // 6 doubles per thread per tile

__global__ void __launch_bounds__(kPeerThreads)
k_map_reduce_cov_mc(PeerArgs a, double *mc, int64_t tile_first, int64_t n_tiles,
                    const double *__restrict__ cov) {
    __shared__ double sh[kPeerTile * 3];
    __shared__ bool is_last;
    const int tid = threadIdx.x;
    const int world = a.world;
    const unsigned long long epoch = *reinterpret_cast<volatile unsigned long long *>(a.epoch_ctr) + 1;

    // ---- start barrier: every rank has finished writing its local map (pass 1) ------------
    if (blockIdx.x == 0 && tid < world) {
        __threadfence_system();
        st_release_sys(a.flags[tid] + a.rank, epoch);
    }
    if (tid < world) {
        const unsigned long long *f = a.flags[a.rank] + tid;
        while (ld_acquire_sys(f) < epoch) {
        }
    }
    __syncthreads();

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t e0 = (tile_first + tile) * (int64_t)(kPeerTile * 3); // double index
        double v[kMcPer];
#pragma unroll
        for (int k = 0; k < kMcPer; ++k) v[k] = mc_ld_reduce_f64(mc + e0 + k * kPeerThreads + tid);
#pragma unroll
        for (int k = 0; k < kMcPer; ++k) sh[k * kPeerThreads + tid] = v[k];
        __syncthreads();
        {
            double *z = sh + 6 * tid; // thread t owns pixels 2t, 2t+1 of the tile
            const double *c = cov + ((tile_first + tile) * (int64_t)kPeerTile + 2 * tid) * 6;
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const double *m = c + 6 * p;
                double v0 = z[3 * p], v1 = z[3 * p + 1], v2 = z[3 * p + 2];
                double t0 = 0.0, t1 = 0.0, t2 = 0.0; // order of toast_map_cov.cpp:509-517
                t0 += __ldg(m) * v0;
                t0 += __ldg(m + 1) * v1;
                t1 += __ldg(m + 1) * v0;
                t0 += __ldg(m + 2) * v2;
                t2 += __ldg(m + 2) * v0;
                t1 += __ldg(m + 3) * v1;
                t1 += __ldg(m + 4) * v2;
                t2 += __ldg(m + 4) * v1;
                t2 += __ldg(m + 5) * v2;
                z[3 * p] = t0;
                z[3 * p + 1] = t1;
                z[3 * p + 2] = t2;
            }
        }
        __syncthreads();
        const double2 *sh2 = reinterpret_cast<const double2 *>(sh);
        double2 *dst = reinterpret_cast<double2 *>(mc + e0);
#pragma unroll
        for (int k = 0; k < kPeerPer; ++k) mc_st_16(dst + k * kPeerThreads + tid, sh2[k * kPeerThreads + tid]);
        __syncthreads();
    }

    // ---- end barrier: every rank's slice has landed in my map ------------------------------
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
        unsigned int done = atomicAdd(a.counter, 1u);
        is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        if (tid < world) {
            __threadfence_system();
            st_release_sys(a.flags[tid] + kMaxPeers + a.rank, epoch);
            const unsigned long long *f = a.flags[a.rank] + kMaxPeers + tid;
            while (ld_acquire_sys(f) < epoch) {
            }
        }
        if (tid == 0) {
            *a.counter = 0u;
            *a.epoch_ctr = epoch;
        }
    }
}

int g_use_multimem = 1; // tb_set_option("multimem", 0/1)

} // namespace

// CTAs per SM of the reduction kernels (tb_set_option("peer_ctas", n)).  Stand-alone the kernel
// wants many resident CTAs (peer-load latency); pipelined with the LHS passes it has to leave
// most of every SM to them.
int tb_peer_ctas_per_sm = 12;

extern "C" {

int tb_peer_set_multimem(int on) {
    g_use_multimem = on ? 1 : 0;
    return 0;
}

// Attach to buffers the caller has already made peer-visible (symmetric memory): `maps` and
// `flags` are world-long arrays of device addresses, rank-ordered (flags: 2 x 16 u64 per rank,
// zeroed); mc_map is the NVLS multicast address of the map buffer or NULL.
tb_peer *tb_peer_attach(int rank, int world, size_t map_bytes, const uint64_t *maps,
                        const uint64_t *flags, uint64_t mc_map) {
    try {
        tbr::require_device();
        TB_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world,
                   "bad rank / world size");
        TB_REQUIRE(maps && flags, "NULL argument");
        tb_peer *p = new tb_peer();
        p->rank = rank;
        p->world = world;
        p->map_bytes = map_bytes;
        p->owned = false;
        for (int q = 0; q < world; ++q) {
            TB_REQUIRE(maps[q] != 0 && flags[q] != 0, "NULL peer address");
            p->peer_map[q] = reinterpret_cast<double *>(maps[q]);
            p->peer_flags[q] = reinterpret_cast<unsigned long long *>(flags[q]);
        }
        p->map = p->peer_map[rank];
        p->flags = p->peer_flags[rank];
        p->mc_map = reinterpret_cast<double *>(mc_map);
        TB_CUDA(cudaMalloc(&p->counter, sizeof(unsigned int)));
        TB_CUDA(cudaMemset(p->counter, 0, sizeof(unsigned int)));
        TB_CUDA(cudaMalloc(&p->epoch_ctr, sizeof(unsigned long long)));
        TB_CUDA(cudaMemset(p->epoch_ctr, 0, sizeof(unsigned long long)));
        p->opened = true;
        return p;
    } catch (const tbr::Error &e) {
        tbr::set_error(e.code, e.msg);
        return nullptr;
    }
}

int tb_peer_has_multicast(const tb_peer *p) { return (p && p->mc_map) ? 1 : 0; }

tb_peer *tb_peer_create(int rank, int world, size_t map_bytes) {
    try {
        tbr::require_device();
        TB_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world,
                   "bad rank / world size");
        tb_peer *p = new tb_peer();
        p->rank = rank;
        p->world = world;
        p->map_bytes = map_bytes;
        TB_CUDA(cudaMalloc(&p->map, map_bytes ? map_bytes : 16));
        TB_CUDA(cudaMemset(p->map, 0, map_bytes ? map_bytes : 16));
        TB_CUDA(cudaMalloc(&p->flags, sizeof(unsigned long long) * 2 * kMaxPeers));
        TB_CUDA(cudaMemset(p->flags, 0, sizeof(unsigned long long) * 2 * kMaxPeers));
        TB_CUDA(cudaMalloc(&p->counter, sizeof(unsigned int)));
        TB_CUDA(cudaMemset(p->counter, 0, sizeof(unsigned int)));
        TB_CUDA(cudaMalloc(&p->epoch_ctr, sizeof(unsigned long long)));
        TB_CUDA(cudaMemset(p->epoch_ctr, 0, sizeof(unsigned long long)));
        p->peer_map[rank] = p->map;
        p->peer_flags[rank] = p->flags;
        if (world == 1) p->opened = true;
        return p;
    } catch (const tbr::Error &e) {
        tbr::set_error(e.code, e.msg);
        return nullptr;
    }
}

// out: 2 x 64 bytes (IPC handles of the map and flag buffers)
int tb_peer_get_handles(tb_peer *p, void *out) {
    TB_API_BEGIN
    TB_REQUIRE(p && out, "NULL argument");
    cudaIpcMemHandle_t h[2];
    TB_CUDA(cudaIpcGetMemHandle(&h[0], p->map));
    TB_CUDA(cudaIpcGetMemHandle(&h[1], p->flags));
    memcpy(out, h, sizeof(h));
    TB_API_END
}

// all: world x 128 bytes, rank-ordered
int tb_peer_open(tb_peer *p, const void *all) {
    TB_API_BEGIN
    TB_REQUIRE(p && all, "NULL argument");
    const cudaIpcMemHandle_t *h = reinterpret_cast<const cudaIpcMemHandle_t *>(all);
    for (int q = 0; q < p->world; ++q) {
        if (q == p->rank) continue;
        void *m = nullptr, *f = nullptr;
        TB_CUDA(cudaIpcOpenMemHandle(&m, h[2 * q], cudaIpcMemLazyEnablePeerAccess));
        TB_CUDA(cudaIpcOpenMemHandle(&f, h[2 * q + 1], cudaIpcMemLazyEnablePeerAccess));
        p->peer_map[q] = (double *)m;
        p->peer_flags[q] = (unsigned long long *)f;
    }
    p->opened = true;
    TB_API_END
}

void *tb_peer_map_ptr(tb_peer *p) { return p ? p->map : nullptr; }

// binned = cov . sum_over_ranks(zmap), result in every rank's map buffer.  nnz = 3.
int tb_map_reduce_cov(tb_peer *p, int64_t n_pix, const double *cov, void *stream) {
    return tb_map_reduce_cov_range(p, 0, n_pix, cov, stream);
}

// The same for the pixel range [pix_first, pix_first + n_pix) only: lets the caller pipeline the
// reduction of one part of the map with the binning of the next (every rank must issue the same
// sequence of calls; calls on one peer object must be stream-ordered).
int tb_map_reduce_cov_range(tb_peer *p, int64_t pix_first, int64_t n_pix, const double *cov,
                            void *stream) {
    TB_API_BEGIN
    tbr::require_device();
    TB_REQUIRE(p && p->opened, "peer buffers are not opened");
    TB_REQUIRE(pix_first >= 0 && n_pix >= 0 && pix_first % kPeerTile == 0 &&
                   n_pix % kPeerTile == 0,
               "pixel range must be aligned to 256 pixels");
    TB_REQUIRE((size_t)(pix_first + n_pix) * 24 <= p->map_bytes, "map buffer too small");
    int64_t tiles = n_pix / kPeerTile;
    int64_t per = (tiles + p->world - 1) / p->world;
    int64_t first = per * p->rank;
    int64_t mine = tiles - first;
    if (mine > per) mine = per;
    if (mine < 0) mine = 0;
    first += pix_first / kPeerTile;
    PeerArgs a;
    for (int q = 0; q < kMaxPeers; ++q) {
        a.map[q] = p->peer_map[q];
        a.flags[q] = p->peer_flags[q];
    }
    a.rank = p->rank;
    a.world = p->world;
    a.epoch_ctr = p->epoch_ctr;
    a.counter = p->counter;
    int64_t grid = mine < 1 ? 1 : mine;
    int64_t cap = (int64_t)tbr::sm_count() * tb_peer_ctas_per_sm;
    if (grid > cap) grid = cap;
    cudaStream_t st = (cudaStream_t)stream;
    if (p->mc_map != nullptr && g_use_multimem && p->world > 1) {
        int64_t gm = mine < 1 ? 1 : mine;
        int64_t capm = (int64_t)tbr::sm_count() * (tb_peer_ctas_per_sm < 8 ? tb_peer_ctas_per_sm : 8);
        if (gm > capm) gm = capm;
        k_map_reduce_cov_mc<<<(unsigned)gm, kPeerThreads, 0, st>>>(a, p->mc_map, first, mine, cov);
        TB_CUDA(cudaGetLastError());
        tbr::count_launch();
        return TB_OK;
    }
    switch (p->world) {
    case 2: k_map_reduce_cov<2><<<(unsigned)grid, kPeerThreads, 0, st>>>(a, first, mine, cov); break;
    case 4: k_map_reduce_cov<4><<<(unsigned)grid, kPeerThreads, 0, st>>>(a, first, mine, cov); break;
    case 8: k_map_reduce_cov<8><<<(unsigned)grid, kPeerThreads, 0, st>>>(a, first, mine, cov); break;
    default: k_map_reduce_cov<0><<<(unsigned)grid, kPeerThreads, 0, st>>>(a, first, mine, cov);
    }
    TB_CUDA(cudaGetLastError());
    tbr::count_launch();
    TB_API_END
}

void tb_peer_destroy(tb_peer *p) {
    if (!p) return;
    if (!p->owned) {
        if (p->counter) cudaFree(p->counter);
        if (p->epoch_ctr) cudaFree(p->epoch_ctr);
        delete p;
        return;
    }
    for (int q = 0; q < p->world; ++q) {
        if (q == p->rank) continue;
        if (p->peer_map[q]) cudaIpcCloseMemHandle(p->peer_map[q]);
        if (p->peer_flags[q]) cudaIpcCloseMemHandle(p->peer_flags[q]);
    }
    if (p->map) cudaFree(p->map);
    if (p->flags) cudaFree(p->flags);
    if (p->counter) cudaFree(p->counter);
    if (p->epoch_ctr) cudaFree(p->epoch_ctr);
    delete p;
}

} // extern "C"
