// tb_prior.cu -- the Offset template's noise prior and its preconditioner on the device.
//
// Reference: templates/offset/offset.py:884-960 (Offset._add_prior) and :962-1010
// (Offset._apply_precond with use_noise_prior=True).  The reference runs both on the host with
// scipy, one (detector, observation, view) segment at a time, and raises NotImplementedError when
// asked to use an accelerator (:888-891, :964-967) -- with a noise prior every PCG iteration
// bounces the amplitude vectors through the host.  Here the filters / factors are uploaded once
// (tb_offset_prior_create) and each application is ONE launch over all segments.
//
//   add_prior   out[seg] += convolve(in[seg], filter[seg], "same");  out[flagged] = 0
//   precond     banded:   out[seg] = cho_solve_banded(factor[seg], in[seg]);  out[flagged] = 0
//               Toeplitz: out[seg] = convolve(in[seg], kernel[seg], "same");  out[flagged] = 0
//   a segment whose detector is cut (start offset < 0) gives zeros.
//
// Parallelisation: the convolution is one thread per output amplitude (tiles of 256 amplitudes
// of one segment per CTA; taps are broadcast loads, the input window is L1-resident).  The banded
// solve is a length-n recurrence per segment: one thread per segment (k_prior_banded) is the
// right shape for ground scans (thousands of short segments per detector set) and the wrong one
// for a single 12-hour satellite view, which the partitioned form (k_pb_*, below) cuts into
// chunks that are solved in parallel.
#include "tb_prior.cuh"
#include "tb_runtime.cuh"

#include <vector>

struct tb_offset_prior {
    int64_t n_amp = 0, n_seg = 0;
    void *blob = nullptr; // one allocation: int64 tables, block table, then doubles
    const int64_t *seg_start = nullptr, *seg_len = nullptr;
    const int64_t *filt_start = nullptr, *filt_len = nullptr;
    const int64_t *prec_start = nullptr, *prec_width = nullptr;
    const int64_t *blocks = nullptr; // [n_blocks][2] {segment, first amplitude of the tile}
    int64_t n_blocks = 0;
    const double *filters = nullptr, *precond = nullptr;
    int precond_mode = 0;
    // partitioned banded solve (tb_set_option("prior_chunk", m) before create; 0 = off)
    bool has_part = false;
    void *part_blob = nullptr;
    tbp::PartView part;
};

// chunk length of the partitioned banded solve for priors created from now on; 0 = one thread
// per segment.  One thread per segment is the wrong shape for long segments (one 12-hour view =
// 43 200 baselines per detector: 402 ms per application on the C4 shard, profiles/README.md);
// the partitioned form with 1024-baseline chunks (12 ms; 256: 20 ms, 4096: 38 ms; local solves per chunk, a (w-1)-wide boundary recurrence, a correction
// with precomputed homogeneous responses) is the default.  Both forms are held to scipy's
// cho_solve_banded on the host (tests/test_offset_prior.py) and on the device
// (tests/test_gpu_prior.py).
int tb_prior_chunk = 1024;

namespace {

constexpr int kConvThreads = 256;
constexpr int kSolveThreads = 64;

// MODE 0: add_prior (accumulate, then zero flagged); MODE 1: Toeplitz precond (overwrite)
template <int MODE>
__global__ void __launch_bounds__(kConvThreads)
k_prior_conv(const int64_t *__restrict__ blocks, const int64_t *__restrict__ seg_start,
             const int64_t *__restrict__ seg_len, const int64_t *__restrict__ f_start,
             const int64_t *__restrict__ f_len, const double *__restrict__ taps,
             const double *__restrict__ in, const uint8_t *__restrict__ flags,
             double *__restrict__ out) {
    const int64_t seg = blocks[2 * (int64_t)blockIdx.x];
    const int64_t i = blocks[2 * (int64_t)blockIdx.x + 1] + threadIdx.x;
    const int64_t n = seg_len[seg];
    if (i >= n) return;
    const int64_t g = seg_start[seg] + i;
    const int64_t fs = f_start[seg];
    double v = 0.0;
    if (fs >= 0 && flags[g] == 0) {
        v = tbp::conv_same_at(in + seg_start[seg], n, taps + fs, f_len[seg], i);
        if (MODE == 0) v += out[g];
    }
    out[g] = v;
}

__global__ void __launch_bounds__(kSolveThreads)
k_prior_banded(int64_t n_seg, const int64_t *__restrict__ seg_start,
               const int64_t *__restrict__ seg_len, const int64_t *__restrict__ p_start,
               const int64_t *__restrict__ p_width, const double *__restrict__ factors,
               const double *__restrict__ in, const uint8_t *__restrict__ flags,
               double *__restrict__ out) {
    const int64_t seg = (int64_t)blockIdx.x * kSolveThreads + threadIdx.x;
    if (seg >= n_seg) return;
    const int64_t n = seg_len[seg], s0 = seg_start[seg], ps = p_start[seg];
    if (ps < 0) {
        for (int64_t j = 0; j < n; ++j) out[s0 + j] = 0.0;
        return;
    }
    tbp::banded_cho_solve(factors + ps, p_width[seg], n, in + s0, out + s0);
    for (int64_t j = 0; j < n; ++j)
        if (flags[s0 + j] != 0) out[s0 + j] = 0.0;
}

// ---- partitioned banded solve: six launches, per-thread code in tb_prior.cuh ---------------------
template <bool FWD>
__global__ void __launch_bounds__(kSolveThreads)
k_pb_chunk(tbp::PartView v, const double *__restrict__ in, double *out) {
    const int64_t g = (int64_t)blockIdx.x * kSolveThreads + threadIdx.x;
    if (g >= v.n_chunk) return;
    if (FWD) tbp::pb_fwd_local(v, g, in, out);
    else tbp::pb_bwd_local(v, g, out);
}

template <bool FWD>
__global__ void __launch_bounds__(kSolveThreads)
k_pb_seg(tbp::PartView v, const double *out) {
    const int64_t seg = (int64_t)blockIdx.x * kSolveThreads + threadIdx.x;
    if (seg >= v.n_seg) return;
    if (FWD) tbp::pb_tails(v, seg, out);
    else tbp::pb_heads(v, seg, out);
}

template <bool FWD>
__global__ void __launch_bounds__(kConvThreads)
k_pb_row(tbp::PartView v, const int64_t *__restrict__ blocks, const uint8_t *__restrict__ flags,
         double *out) {
    const int64_t seg = blocks[2 * (int64_t)blockIdx.x];
    const int64_t j = blocks[2 * (int64_t)blockIdx.x + 1] + threadIdx.x;
    if (j >= v.seg_len[seg]) return;
    if (FWD) tbp::pb_fwd_correct(v, seg, j, out);
    else tbp::pb_bwd_correct(v, seg, j, flags, out);
}

void build_partition(tb_offset_prior *p, const tb_offset_prior_desc *d, int64_t chunk) {
    tbp::PartTables T;
    tbp::build_part_tables(d->n_seg, d->seg_len, d->prec_start, d->prec_width, d->precond, chunk, T);
    const size_t ni = T.seg_chunk0.size() + T.seg_m.size() + T.chunk_seg.size() + T.g_off.size();
    const size_t nd = T.Gf.size() + T.Gb.size() + 2 * (size_t)(T.n_chunk * T.qmax + 1);
    TB_CUDA(cudaMalloc(&p->part_blob, ni * sizeof(int64_t) + nd * sizeof(double)));
    int64_t *di = (int64_t *)p->part_blob;
    double *dd = (double *)(di + ni);
    auto up_i = [&](const std::vector<int64_t> &v, int64_t *&dst) {
        const int64_t *at = dst;
        if (!v.empty())
            TB_CUDA(cudaMemcpy(dst, v.data(), v.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
        dst += v.size();
        return at;
    };
    tbp::PartView &v = p->part;
    v.n_seg = d->n_seg;
    v.n_chunk = T.n_chunk;
    v.qmax = T.qmax;
    v.seg_start = p->seg_start;
    v.seg_len = p->seg_len;
    v.p_start = p->prec_start;
    v.p_width = p->prec_width;
    v.factors = p->precond;
    v.seg_chunk0 = up_i(T.seg_chunk0, di);
    v.seg_m = up_i(T.seg_m, di);
    v.chunk_seg = up_i(T.chunk_seg, di);
    v.g_off = up_i(T.g_off, di);
    TB_CUDA(cudaMemcpy(dd, T.Gf.data(), T.Gf.size() * sizeof(double), cudaMemcpyHostToDevice));
    v.Gf = dd;
    dd += T.Gf.size();
    TB_CUDA(cudaMemcpy(dd, T.Gb.data(), T.Gb.size() * sizeof(double), cudaMemcpyHostToDevice));
    v.Gb = dd;
    dd += T.Gb.size();
    v.tails = dd;
    v.heads = dd + (T.n_chunk * T.qmax + 1);
    TB_CUDA(cudaMemset(v.tails, 0, 2 * (size_t)(T.n_chunk * T.qmax + 1) * sizeof(double)));
    p->has_part = true;
}

} // namespace

extern "C" {

tb_offset_prior *tb_offset_prior_create(const tb_offset_prior_desc *d) {
    try {
        tbr::require_device();
        TB_REQUIRE(d != nullptr, "NULL descriptor");
        TB_REQUIRE(d->n_amp >= 0 && d->n_seg >= 0, "bad sizes");
        TB_REQUIRE(d->n_seg == 0 || (d->seg_start && d->seg_len && d->filt_start && d->filt_len),
                   "missing segment tables");
        TB_REQUIRE(d->precond_mode == TB_PRECOND_TOEPLITZ || d->precond_mode == TB_PRECOND_BANDED,
                   "precond_mode must be TB_PRECOND_TOEPLITZ or TB_PRECOND_BANDED");
        TB_REQUIRE(d->n_seg == 0 || (d->prec_start && d->prec_width), "missing preconditioner tables");
        const int64_t ns = d->n_seg;
        std::vector<int64_t> blocks;
        int64_t expect = 0;
        for (int64_t s = 0; s < ns; ++s) {
            const int64_t n = d->seg_len[s], s0 = d->seg_start[s];
            TB_REQUIRE(n >= 0 && s0 >= 0 && s0 + n <= d->n_amp, "segment outside the amplitudes");
            TB_REQUIRE(s0 >= expect, "segments must be disjoint and in increasing order");
            expect = s0 + n;
            if (d->filt_start[s] >= 0) {
                TB_REQUIRE(d->filt_len[s] >= 1 &&
                               d->filt_start[s] + d->filt_len[s] <= d->n_filter_values,
                           "filter outside the filter array");
            }
            if (d->prec_start[s] >= 0) {
                int64_t need = d->precond_mode == TB_PRECOND_BANDED ? d->prec_width[s] * n
                                                                    : d->prec_width[s];
                TB_REQUIRE(d->prec_width[s] >= 1 &&
                               d->prec_start[s] + need <= d->n_precond_values,
                           "preconditioner outside its array");
            }
            for (int64_t i0 = 0; i0 < n; i0 += kConvThreads) {
                blocks.push_back(s);
                blocks.push_back(i0);
            }
        }
        tb_offset_prior *p = new tb_offset_prior();
        p->n_amp = d->n_amp;
        p->n_seg = ns;
        p->precond_mode = d->precond_mode;
        p->n_blocks = (int64_t)blocks.size() / 2;
        const size_t n_i = (size_t)(6 * ns) + blocks.size();
        const size_t n_d = (size_t)d->n_filter_values + (size_t)d->n_precond_values;
        std::vector<int64_t> ib(n_i > 0 ? n_i : 1, 0);
        for (int64_t s = 0; s < ns; ++s) {
            ib[s] = d->seg_start[s];
            ib[ns + s] = d->seg_len[s];
            ib[2 * ns + s] = d->filt_start[s];
            ib[3 * ns + s] = d->filt_len[s];
            ib[4 * ns + s] = d->prec_start[s];
            ib[5 * ns + s] = d->prec_width[s];
        }
        for (size_t k = 0; k < blocks.size(); ++k) ib[6 * ns + k] = blocks[k];
        const size_t ibytes = ib.size() * sizeof(int64_t);
        TB_CUDA(cudaMalloc(&p->blob, ibytes + (n_d > 0 ? n_d : 1) * sizeof(double)));
        TB_CUDA(cudaMemcpy(p->blob, ib.data(), ibytes, cudaMemcpyHostToDevice));
        double *dd = (double *)((char *)p->blob + ibytes);
        if (d->n_filter_values > 0)
            TB_CUDA(cudaMemcpy(dd, d->filters, sizeof(double) * d->n_filter_values,
                               cudaMemcpyHostToDevice));
        if (d->n_precond_values > 0)
            TB_CUDA(cudaMemcpy(dd + d->n_filter_values, d->precond,
                               sizeof(double) * d->n_precond_values, cudaMemcpyHostToDevice));
        const int64_t *di = (const int64_t *)p->blob;
        p->seg_start = di;
        p->seg_len = di + ns;
        p->filt_start = di + 2 * ns;
        p->filt_len = di + 3 * ns;
        p->prec_start = di + 4 * ns;
        p->prec_width = di + 5 * ns;
        p->blocks = di + 6 * ns;
        p->filters = dd;
        p->precond = dd + d->n_filter_values;
        if (tb_prior_chunk > 0 && d->precond_mode == TB_PRECOND_BANDED && ns > 0)
            build_partition(p, d, tb_prior_chunk);
        return p;
    } catch (const tbr::Error &e) {
        tbr::set_error(e.code, e.msg);
        return nullptr;
    }
}

void tb_offset_prior_destroy(tb_offset_prior *p) {
    if (p == nullptr) return;
    if (p->blob) cudaFree(p->blob);
    if (p->part_blob) cudaFree(p->part_blob);
    delete p;
}

int tb_offset_prior_add(const tb_offset_prior *p, const double *amplitudes_in,
                        const uint8_t *amplitude_flags, double *amplitudes_out, int mem,
                        void *stream) {
    TB_API_BEGIN
    tbr::require_device();
    TB_REQUIRE(p && amplitudes_in && amplitude_flags && amplitudes_out, "NULL argument");
    TB_REQUIRE(amplitudes_in != amplitudes_out, "add_prior cannot run in place");
    tbr::Resolver R(mem, stream);
    const double *d_in = R.in(amplitudes_in, p->n_amp);
    const uint8_t *d_f = R.in(amplitude_flags, p->n_amp);
    double *d_out = R.inout(amplitudes_out, p->n_amp);
    if (p->n_blocks > 0) {
        TB_REQUIRE(p->n_blocks < 2147483647LL, "grid too large");
        k_prior_conv<0><<<(unsigned)p->n_blocks, kConvThreads, 0, R.stream()>>>(
            p->blocks, p->seg_start, p->seg_len, p->filt_start, p->filt_len, p->filters, d_in, d_f,
            d_out);
        TB_CUDA(cudaGetLastError());
        tbr::count_launch();
    }
    R.finish();
    TB_API_END
}

int tb_offset_prior_precond(const tb_offset_prior *p, const double *amplitudes_in,
                            const uint8_t *amplitude_flags, double *amplitudes_out, int mem,
                            void *stream) {
    TB_API_BEGIN
    tbr::require_device();
    TB_REQUIRE(p && amplitudes_in && amplitude_flags && amplitudes_out, "NULL argument");
    TB_REQUIRE(amplitudes_in != amplitudes_out, "the preconditioner cannot run in place");
    tbr::Resolver R(mem, stream);
    const double *d_in = R.in(amplitudes_in, p->n_amp);
    const uint8_t *d_f = R.in(amplitude_flags, p->n_amp);
    double *d_out = R.out(amplitudes_out, p->n_amp);
    if (p->precond_mode == TB_PRECOND_TOEPLITZ) {
        if (p->n_blocks > 0) {
            TB_REQUIRE(p->n_blocks < 2147483647LL, "grid too large");
            k_prior_conv<1><<<(unsigned)p->n_blocks, kConvThreads, 0, R.stream()>>>(
                p->blocks, p->seg_start, p->seg_len, p->prec_start, p->prec_width, p->precond, d_in,
                d_f, d_out);
            TB_CUDA(cudaGetLastError());
            tbr::count_launch();
        }
    } else if (p->has_part) {
        const tbp::PartView &v = p->part;
        const unsigned nbc = (unsigned)((v.n_chunk + kSolveThreads - 1) / kSolveThreads);
        const unsigned nbs = (unsigned)((v.n_seg + kSolveThreads - 1) / kSolveThreads);
        TB_REQUIRE(p->n_blocks < 2147483647LL, "grid too large");
        const unsigned nbr = (unsigned)p->n_blocks;
        cudaStream_t st = R.stream();
        if (v.n_chunk > 0 && nbr > 0) {
            k_pb_chunk<true><<<nbc, kSolveThreads, 0, st>>>(v, d_in, d_out);
            k_pb_seg<true><<<nbs, kSolveThreads, 0, st>>>(v, d_out);
            k_pb_row<true><<<nbr, kConvThreads, 0, st>>>(v, p->blocks, d_f, d_out);
            k_pb_chunk<false><<<nbc, kSolveThreads, 0, st>>>(v, d_in, d_out);
            k_pb_seg<false><<<nbs, kSolveThreads, 0, st>>>(v, d_out);
            k_pb_row<false><<<nbr, kConvThreads, 0, st>>>(v, p->blocks, d_f, d_out);
            TB_CUDA(cudaGetLastError());
            tbr::count_launch(6);
        }
    } else if (p->n_seg > 0) {
        const int64_t nb = (p->n_seg + kSolveThreads - 1) / kSolveThreads;
        k_prior_banded<<<(unsigned)nb, kSolveThreads, 0, R.stream()>>>(
            p->n_seg, p->seg_start, p->seg_len, p->prec_start, p->prec_width, p->precond, d_in, d_f,
            d_out);
        TB_CUDA(cudaGetLastError());
        tbr::count_launch();
    }
    R.finish();
    TB_API_END
}

} // extern "C"
