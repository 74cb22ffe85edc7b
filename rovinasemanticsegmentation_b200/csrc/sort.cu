// Radix sort of (64-bit key, 32-bit value) pairs for the host-side orchestration code (crf.cu): CUB's DeviceRadixSort,
// which ships with the CUDA toolkit - library code, kept out of the kernel files so that they do not pull in CUB.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "kernels.hpp"

namespace rss {

// sorts n pairs by the low `bits` bits of the keys; tmp is a grow-only scratch buffer owned by the caller
cudaError_t sort_pairs_u64(cudaStream_t st, DevBuf& tmp, const uint64_t* keys_in, uint64_t* keys_out, const uint32_t* vals_in,
                           uint32_t* vals_out, size_t n, int bits) {
    if (n >= (1ull << 31)) return cudaErrorInvalidValue;
    size_t need = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, need, keys_in, keys_out, vals_in, vals_out, (int)n, 0, bits, st);
    if (e != cudaSuccess) return e;
    e = tmp.reserve(need);
    if (e != cudaSuccess) return e;
    size_t have = tmp.cap;
    return cub::DeviceRadixSort::SortPairs(tmp.ptr, have, keys_in, keys_out, vals_in, vals_out, (int)n, 0, bits, st);
}

}  // namespace rss
