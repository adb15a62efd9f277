// Row gather: dst[i, :] = src[idx[i], :].  Used for (a) ring (time-major) -> flat dataset order
// and (b) dataset -> minibatch rows under the epoch permutation.  HBM-bound: each row is read
// once and written once; 128-bit accesses whenever the row width allows.
#include "common.cuh"

namespace ppoaf {

template <typename IdxT, typename VecT>
__global__ void gather_rows_vec_kernel(const VecT* __restrict__ src, int64_t src_stride_vecs,
                                       const IdxT* __restrict__ idx, VecT* __restrict__ dst, int64_t n_rows,
                                       int32_t vecs_per_row) {
    const int64_t total = n_rows * vecs_per_row;
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t g = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; g < total; g += stride) {
        const int64_t row = g / vecs_per_row;
        const int32_t v = int32_t(g - row * vecs_per_row);
        const int64_t s = int64_t(idx[row]);
        dst[g] = src[s * src_stride_vecs + v];
    }
}

}  // namespace ppoaf

using namespace ppoaf;

template <typename IdxT>
static int launch_gather(const void* src, int64_t src_stride, const void* idx, void* dst, int64_t n_rows,
                         int64_t row_bytes, cudaStream_t s) {
    const uintptr_t al = reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | uintptr_t(row_bytes) |
                         uintptr_t(src_stride);
    const int threads = 256;
    const int64_t cap = int64_t(sm_count()) * 8;
    auto grid_for = [&](int64_t total) {
        int64_t b = ceil_div64(total, threads);
        return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
    };
    if (al % 16 == 0) {
        const int32_t vpr = int32_t(row_bytes / 16);
        gather_rows_vec_kernel<IdxT, int4><<<grid_for(n_rows * vpr), threads, 0, s>>>(
            (const int4*)src, src_stride / 16, (const IdxT*)idx, (int4*)dst, n_rows, vpr);
    } else if (al % 8 == 0) {
        const int32_t vpr = int32_t(row_bytes / 8);
        gather_rows_vec_kernel<IdxT, int2><<<grid_for(n_rows * vpr), threads, 0, s>>>(
            (const int2*)src, src_stride / 8, (const IdxT*)idx, (int2*)dst, n_rows, vpr);
    } else if (al % 4 == 0) {
        const int32_t vpr = int32_t(row_bytes / 4);
        gather_rows_vec_kernel<IdxT, int32_t><<<grid_for(n_rows * vpr), threads, 0, s>>>(
            (const int32_t*)src, src_stride / 4, (const IdxT*)idx, (int32_t*)dst, n_rows, vpr);
    } else {
        const int32_t vpr = int32_t(row_bytes);
        gather_rows_vec_kernel<IdxT, uint8_t><<<grid_for(n_rows * vpr), threads, 0, s>>>(
            (const uint8_t*)src, src_stride / 1, (const IdxT*)idx, (uint8_t*)dst, n_rows, vpr);
    }
    PPOAF_CHECK_LAUNCH("ppoaf_gather_rows");
    return 0;
}

extern "C" int ppoaf_gather_rows(const void* src, int64_t src_stride_bytes, const void* idx, int idx_is_64, void* dst,
                                 int64_t n_rows, int64_t row_bytes, void* stream) {
    PPOAF_CHECK_ARG(n_rows >= 0 && row_bytes > 0 && row_bytes < (int64_t(1) << 31), "ppoaf_gather_rows: bad sizes");
    if (src_stride_bytes == 0) src_stride_bytes = row_bytes;
    PPOAF_CHECK_ARG(src_stride_bytes >= row_bytes, "ppoaf_gather_rows: source stride smaller than the row");
    if (n_rows == 0) return 0;
    if (idx_is_64) return launch_gather<int64_t>(src, src_stride_bytes, idx, dst, n_rows, row_bytes, (cudaStream_t)stream);
    return launch_gather<int32_t>(src, src_stride_bytes, idx, dst, n_rows, row_bytes, (cudaStream_t)stream);
}
