// Standalone bring-up test for tcgen05.mma kind::tf32 with NO-swizzle canonical smem layouts,
// both operand majors, and the 3xTF32 split.  nvcc -gencode arch=compute_100a,code=sm_100a umma_test.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

constexpr int M = 128, N = 64, K = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 0) {
    uint64_t d = (uint64_t)layout_type << 61;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
    return d;                // layout_type = 0 (no swizzle), base_offset = 0, lbo_mode = 0
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ float to_tf32_rna(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

// mode bit0: A is MN-major in global ([k][m]); bit1: B is MN-major ([k][n]); bit2: 3xTF32
__global__ void __launch_bounds__(128) umma_test_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                        float* __restrict__ D, int mode) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar;
    const bool a_mn = mode & 1, b_mn = mode & 2, split = mode & 4;
    float* Ahi = reinterpret_cast<float*>(smem);                 // M*K floats
    float* Alo = Ahi + M * K;
    float* Bhi = Alo + M * K;                                    // N*K floats
    float* Blo = Bhi + N * K;
    const int tid = threadIdx.x, warp = tid >> 5;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    // ---- fill smem: per 32-wide K chunk one tile; K-major = SWIZZLE_128B rows, MN-major = SWIZZLE_128B_BASE32B atoms ----
    //  K-major  operand (rows R): chunk (row, kq)  -> [kq][row][4]            LBO = R*16, SBO = 128
    //  MN-major operand (rows R): chunk (k, rq)    -> [k/8][rq][k%8][4]       SBO = 128, K-group stride = (R/4)*128
    for (int c = tid; c < M * K / 4; c += 128) {
        float4 v;
        int dst;
        if (!a_mn) { const int row = c / (K / 4), kq = c % (K / 4); v = *reinterpret_cast<const float4*>(A + row * K + kq * 4); dst = (kq / 8) * (M * 32) + row * 32 + (((kq % 8) ^ (row & 7)) * 4); }
        else       { const int k = c / (M / 4), rq = c % (M / 4);   v = *reinterpret_cast<const float4*>(A + k * M + rq * 4);  dst = (k / 32) * (M * 32) + (rq / 8) * 1024 + ((k % 32) / 4) * 128 + (k % 4) * 32 + ((((rq / 2) % 4) ^ (k % 4)) * 8) + (rq % 2) * 4; }
        float h[4] = {v.x, v.y, v.z, v.w}, l[4];
        for (int j = 0; j < 4; ++j) { if (split) { const float hi = to_tf32_rna(h[j]); l[j] = to_tf32_rna(h[j] - hi); h[j] = hi; } else l[j] = 0.f; }
        *reinterpret_cast<float4*>(Ahi + dst) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(Alo + dst) = make_float4(l[0], l[1], l[2], l[3]);
    }
    for (int c = tid; c < N * K / 4; c += 128) {
        float4 v;
        int dst;
        if (!b_mn) { const int row = c / (K / 4), kq = c % (K / 4); v = *reinterpret_cast<const float4*>(B + row * K + kq * 4); dst = (kq / 8) * (N * 32) + row * 32 + (((kq % 8) ^ (row & 7)) * 4); }
        else       { const int k = c / (N / 4), rq = c % (N / 4);   v = *reinterpret_cast<const float4*>(B + k * N + rq * 4);  dst = (k / 32) * (N * 32) + (rq / 8) * 1024 + ((k % 32) / 4) * 128 + (k % 4) * 32 + ((((rq / 2) % 4) ^ (k % 4)) * 8) + (rq % 2) * 4; }
        float h[4] = {v.x, v.y, v.z, v.w}, l[4];
        for (int j = 0; j < 4; ++j) { if (split) { const float hi = to_tf32_rna(h[j]); l[j] = to_tf32_rna(h[j] - hi); h[j] = hi; } else l[j] = 0.f; }
        *reinterpret_cast<float4*>(Bhi + dst) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(Blo + dst) = make_float4(l[0], l[1], l[2], l[3]);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
                               ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        uint32_t acc = 0;
        const int n_terms = split ? 3 : 1;
        for (int term = 0; term < n_terms; ++term) {
            const float* Ap = (term == 2) ? Alo : Ahi;           // hi*hi, hi*lo, lo*hi
            const float* Bp = (term == 1) ? Blo : Bhi;
            for (int s = 0; s < K / 8; ++s) {
                uint64_t da, db;
                if (!a_mn) da = make_desc(smem_u32(Ap) + (s / 4) * (M * 128) + (s % 4) * 32, 0, 1024, 2);
                else       da = make_desc(smem_u32(Ap) + (s / 4) * (M * 128) + (s % 4) * 1024, 4096, 512, 1);
                if (!b_mn) db = make_desc(smem_u32(Bp) + (s / 4) * (N * 128) + (s % 4) * 32, 0, 1024, 2);
                else       db = make_desc(smem_u32(Bp) + (s / 4) * (N * 128) + (s % 4) * 1024, 4096, 512, 1);
                umma_tf32(tmem, da, db, idesc, acc);
                acc = 1;
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar)) : "memory");
    }
    // ---- wait for the accumulator, read it back (warp w owns TMEM lanes 32w..32w+31) ----
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}\n"
                : "=r"(done) : "r"(smem_u32(&s_bar)), "r"(0) : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
            "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) D[tid * N + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64));
}

int main() {
    std::vector<float> hA(M * K), hB(N * K), hD(M * N);
    srand(1);
    for (auto& x : hA) x = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    for (auto& x : hB) x = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    float *dA, *dB, *dD;
    cudaMalloc(&dA, hA.size() * 4); cudaMalloc(&dB, hB.size() * 4); cudaMalloc(&dD, hD.size() * 4);
    const size_t smem = (size_t)(2 * M * K + 2 * N * K) * 4 + 1024;
    cudaFuncSetAttribute(umma_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int bad = 0;
    for (int mode = 0; mode < 8; ++mode) {
        // logical A[m][k], B[n][k]; physical layout depends on the major bits
        std::vector<float> pA(M * K), pB(N * K);
        for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) pA[(mode & 1) ? k * M + m : m * K + k] = hA[m * K + k];
        for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) pB[(mode & 2) ? k * N + n : n * K + k] = hB[n * K + k];
        cudaMemcpy(dA, pA.data(), pA.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, pB.data(), pB.size() * 4, cudaMemcpyHostToDevice);
        cudaMemset(dD, 0, hD.size() * 4);
        umma_test_kernel<<<1, 128, smem>>>(dA, dB, dD, mode);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0, maxref = 0;
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += (double)hA[m * K + k] * hB[n * K + k];
            maxerr = fmax(maxerr, fabs(ref - hD[m * N + n]));
            maxref = fmax(maxref, fabs(ref));
        }
        const double tol = (mode & 4) ? 2e-5 : 2e-2;
        printf("mode %d (A %s, B %s, %s): max abs err %.3e (max |ref| %.2f) %s\n", mode, (mode & 1) ? "MN" : "K ",
               (mode & 2) ? "MN" : "K ", (mode & 4) ? "3xTF32" : "1xTF32", maxerr, maxref, maxerr < tol ? "ok" : "BAD");
        bad += !(maxerr < tol);
    }
    printf(bad ? "UMMA_TEST FAIL\n" : "UMMA_TEST PASS\n");
    return bad;
}
