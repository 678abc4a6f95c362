// tb_sort.cu -- setup-time helper: stable radix sort of (int32 key, int32 value) pairs on the
// device (CUB, shipped with the CUDA toolkit).  Used ONCE per solve by tb_obs_pack_pointing to
// order the crossing list by pixel; nothing on the per-iteration path calls it.
#include <cub/device/device_radix_sort.cuh>

#include "tb_runtime.cuh"

namespace tbr {

void sort_pairs_i32(const int32_t *keys_in, int32_t *keys_out, const int32_t *vals_in,
                    int32_t *vals_out, int64_t n, int end_bit, cudaStream_t st) {
    TB_REQUIRE(n < 2147483647LL, "too many pairs to sort");
    size_t tmp_bytes = 0;
    TB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in, keys_out, vals_in,
                                            vals_out, (int)n, 0, end_bit, st));
    void *tmp = nullptr;
    TB_CUDA(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16));
    cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_in, keys_out, vals_in,
                                                    vals_out, (int)n, 0, end_bit, st);
    cudaError_t e2 = cudaStreamSynchronize(st);
    cudaFree(tmp);
    TB_CUDA(e);
    TB_CUDA(e2);
}

} // namespace tbr
