// Probe: which smem element does tcgen05.mma (tf32, no swizzle) read for A(m,k) / B(n,k) in MN-major mode?
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>
constexpr int M = 128, N = 16, K = 8;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t lt = 0) {
    return ((uint64_t)lt << 61) | (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
// which: 0 = probe A as MN-major (B = K-major identity), 1 = probe B as MN-major (A = K-major identity on rows 0..7)
__global__ void __launch_bounds__(128) probe(float* D, int which, uint32_t lbo, uint32_t sbo, uint32_t lt) {
    __shared__ __align__(1024) float sA[4096];
    __shared__ __align__(1024) float sB[4096];
    __shared__ uint32_t s_tmem; __shared__ __align__(8) uint64_t s_bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(32));
                     asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
    for (int i = tid; i < 4096; i += 128) { sA[i] = 0.f; sB[i] = 0.f; }
    __syncthreads();
    if (which == 0) {
        for (int i = tid; i < 2048; i += 128) sA[i] = float(i);                       // pattern: value = float index
        if (tid < 8) sB[(tid / 4 * N + tid) * 4 + tid % 4] = 1.f;                        // K-major B: [kq][n][4]: B[n=tid][k=tid] = 1
    } else {
        for (int i = tid; i < 2048; i += 128) sB[i] = float(i);
        if (tid < 8) sA[(tid / 4 * M + tid) * 4 + tid % 4] = 1.f;                        // K-major A: A[m=tid][k=tid] = 1
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((which == 0 ? 1u : 0u) << 15) | ((which == 1 ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        uint64_t da = which == 0 ? make_desc(smem_u32(sA), lbo, sbo, lt) : make_desc(smem_u32(sA), M * 16, 128);
        uint64_t db = which == 1 ? make_desc(smem_u32(sB), lbo, sbo, lt) : make_desc(smem_u32(sB), N * 16, 128);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(0) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar)) : "memory");
    }
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&s_bar)), "r"(0) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[16];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) D[tid * 16 + j] = __uint_as_float(r[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32));
}
int main() {
    float* dD; cudaMalloc(&dD, 128 * 16 * 4);
    std::vector<float> h(128 * 16);
    struct { uint32_t lbo, sbo, lt; } cfgs[] = {{1024, 512, 1}, {512, 1024, 1}, {1024, 512, 2}};
    for (int which = 0; which < 2; ++which) for (auto c : cfgs) {
        cudaMemset(dD, 0, 128 * 16 * 4);
        probe<<<1, 128>>>(dD, which, c.lbo, c.sbo, c.lt);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h.data(), dD, h.size() * 4, cudaMemcpyDeviceToHost);
        printf("probe %s MN-major, LBO=%u SBO=%u layout_type=%u\n", which == 0 ? "A" : "B", c.lbo, c.sbo, c.lt);
        if (which == 0) {   // D[m][n=k] = smem float index HW used for A(m,k)
            for (int m : {0, 1, 7, 8, 9, 16, 24, 31, 32, 33, 64, 127}) { printf("  m=%3d:", m); for (int k = 0; k < 8; ++k) printf(" %6.0f", h[m * 16 + k]); printf("\n"); }
        } else {            // D[m=k][n] = smem float index HW used for B(n,k)
            for (int n = 0; n < 16; ++n) { printf("  n=%3d:", n); for (int k = 0; k < 8; ++k) printf(" %6.0f", h[k * 16 + n]); printf("\n"); }
        }
    }
    return 0;
}
