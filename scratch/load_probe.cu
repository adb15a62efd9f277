// How fast can one CTA per SM pull a (32+64) x K fp32 operand panel from L2 into shared memory?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o load_probe load_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int M = 512, N = 256, K = 256;
constexpr int ROWS = 96, PAD = 4;
__device__ long long g_stamp[8];
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// mode 0: LDGSTS.128, lanes along k (coalesced rows); 1: bulk 256 B segments; 2: bulk whole rows (1 KB); 3: LDGSTS 16 lines/instr
// 4: bulk 128 B segments
__global__ void __launch_bounds__(256) probe(const float* __restrict__ X, const float* __restrict__ W, float* out, int mode) {
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x;
    const int m0 = (blockIdx.x / 4) * 32, n0 = (blockIdx.x % 4) * 64;
    const long long t0 = clock64();
    if (mode != 0 && mode != 3) {
        if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
        __syncthreads();
    }
    if (tid == 0) g_stamp[0] = clock64() - t0;
    auto rowp = [&](int r) { return r < 32 ? X + (size_t)(m0 + r) * K : W + (size_t)(n0 + r - 32) * K; };
    if (mode >= 5) {
        // the GEMM kernel's warp-private mapping: warp g owns k in [32 g, 32 g + 32); lanes: kq = lane & 7, row = lane / 8 + 4 j
        const int g = tid >> 5, lane = tid & 31, kq = lane & 7;
        const int blk = mode == 5 ? 128 : 144;          // floats per A block (mode 6: +16 floats of skew)
        const int blkB = mode == 5 ? 256 : 272;
        const int slotf = 2 * blk + 2 * blkB;
        float* ring = smem + g * 4 * slotf;
        for (int j = 0; j < 24; ++j) {
            int row = (lane >> 3) + 4 * j;
            float* dst = ring + (kq >> 1) * slotf + (row < 32 ? (kq & 1) * blk + row * 4 : 2 * blk + (kq & 1) * blkB + (row - 32) * 4);
            uint32_t d = smem_u32(dst);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(rowp(row) + g * 32 + kq * 4) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (tid == 0) g_stamp[1] = clock64() - t0;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
    } else if (mode == 0 || mode == 3) {
        // 96 rows x 64 chunks of 16 B = 6144 chunks / 256 threads = 24 per thread
        for (int j = 0; j < 24; ++j) {
            int c = tid + 256 * j, r, q;
            if (mode == 0) { r = c / 64; q = c % 64; } else { r = c % 96; q = c / 96; }
            uint32_t d = smem_u32(smem + r * (K + PAD) + q * 4);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(rowp(r) + q * 4) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (tid == 0) g_stamp[1] = clock64() - t0;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
    } else {
        const int seg = mode == 1 ? 256 : (mode == 2 ? 1024 : 128);   // bytes
        const int per_row = 1024 / seg, total = ROWS * per_row;
        if (tid == 0) mbar_expect(&bar, ROWS * 1024);
        __syncthreads();
        for (int c = tid; c < total; c += 256) {
            int r = c / per_row, q = c % per_row;
            bulk_g2s(smem + r * (K + PAD) + q * (seg / 4), rowp(r) + q * (seg / 4), seg, &bar);
        }
        if (tid == 0) g_stamp[1] = clock64() - t0;
        mbar_wait(&bar, 0);
    }
    if (tid == 0 && blockIdx.x == 0) g_stamp[2] = clock64() - t0;
    // consume something so nothing is optimised away
    float s = 0.f;
    for (int i = tid; i < ROWS * (K + PAD); i += 256 * 37) s += smem[i];
    if (s == 12345.678f) out[0] = s;
}
int main() {
    float *X, *W, *out;
    cudaMalloc(&X, M * K * 4); cudaMalloc(&W, N * K * 4); cudaMalloc(&out, 4);
    cudaMemset(X, 0, M * K * 4); cudaMemset(W, 0, N * K * 4);
    const int smem = 8 * 4 * (2 * 144 + 2 * 272) * 4;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 7; ++mode) {
        for (int grid : {64, 128}) {
            for (int i = 0; i < 5; ++i) probe<<<grid, 256, smem>>>(X, W, out, mode);
            cudaEventRecord(e0);
            for (int i = 0; i < 200; ++i) probe<<<grid, 256, smem>>>(X, W, out, mode);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            long long st[8]; cudaMemcpyFromSymbol(st, g_stamp, sizeof(st));
            printf("mode %d grid %3d: %.2f us/launch  issue done %lld clk, landed %lld clk (%s)\n", mode, grid, ms * 1e3 / 200, st[1], st[2]); printf("      setup %lld\n", st[0], cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
