// What costs ~2400 clk at the start of a small kernel: cold instruction fetch or first-touch address translation?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int K = 256, PAD = 4, ROWS = 96;
__device__ long long g_stamp[8];
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(256) probe(const float* __restrict__ X, const float* __restrict__ W, float* out, int pretouch) {
    extern __shared__ __align__(128) float smem[];
    const int tid = threadIdx.x;
    const int m0 = (blockIdx.x / 4) * 32, n0 = (blockIdx.x % 4) * 64;
    const long long t0 = clock64();
    float pre = 0.f;
    if (pretouch) {   // translate every page the panel touches before the timed passes
        pre = X[(size_t)(m0 + (tid & 31)) * K] + W[(size_t)(n0 + (tid & 63)) * K];
        if (pre == 123.f) out[1] = pre;
    }
    __syncthreads();
    if (tid == 0) g_stamp[0] = clock64() - t0;
    auto rowp = [&](int r) { return r < 32 ? X + (size_t)(m0 + r) * K : W + (size_t)(n0 + r - 32) * K; };
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
#pragma unroll 4
        for (int j = 0; j < 24; ++j) {
            int c = tid + 256 * j, r = c / 64, q = c % 64;
            uint32_t d = smem_u32(smem + r * (K + PAD) + q * 4);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(rowp(r) + q * 4) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (tid == 0) g_stamp[1 + 2 * pass] = clock64() - t0;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        if (tid == 0) g_stamp[2 + 2 * pass] = clock64() - t0;
    }
    float s = 0.f;
    for (int i = tid; i < ROWS * (K + PAD); i += 256 * 37) s += smem[i];
    if (s == 12345.678f) out[0] = s;
}
int main() {
    float *X, *W, *out;
    cudaMalloc(&X, 512 * K * 4); cudaMalloc(&W, 256 * K * 4); cudaMalloc(&out, 8);
    cudaMemset(X, 0, 512 * K * 4); cudaMemset(W, 0, 256 * K * 4);
    const int smem = ROWS * (K + PAD) * 4;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int pt = 0; pt < 2; ++pt) {
        for (int i = 0; i < 50; ++i) probe<<<128, 256, smem>>>(X, W, out, pt);
        cudaDeviceSynchronize();
        long long st[8]; cudaMemcpyFromSymbol(st, g_stamp, sizeof(st));
        printf("pretouch %d: setup %lld | pass0 issue %lld landed %lld | pass1 issue %lld landed %lld | pass2 issue %lld landed %lld\n",
               pt, st[0], st[1], st[2], st[3], st[4], st[5], st[6]);
    }
    return 0;
}
