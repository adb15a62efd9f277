// fp32 tiled GEMM core for the actor/critic MLP layers at reference minibatch sizes (B = 128..512,
// hidden 64..512), where a layer is far too small to be a dense tensor-core contraction (SURVEY.md
// §8d): FFMA register tiles fed from double-buffered shared memory, 128-bit global loads, fused
// bias / activation / activation-derivative epilogues and a fused row gather on the operand that
// is indexed by sample (the minibatch permutation), so the gathered minibatch never exists in HBM.
//
//   forward      Y[m,n]  = act(sum_k X[idx[m],k] W[n,k] + b[n])        FeedForwardNetwork.forward
//   backward-x   dX[m,k] = (sum_n dZ[m,n] W[n,k]) * act'(Xact[m,k])    autograd of the same
//   backward-w   dW[n,k] = sum_m dZ[m,n] X[idx[m],k] ; db[n] = sum_m dZ[m,n]
#pragma once
#include "common.cuh"

namespace ppoaf {

__device__ __forceinline__ float act_fwd(float x, int act) {
    switch (act) {
        case PPOAF_ACT_RELU: return x > 0.f ? x : 0.f;
        case PPOAF_ACT_LEAKY_RELU: return x > 0.f ? x : 0.01f * x;
        case PPOAF_ACT_TANH: return tanhf(x);
        default: return x;
    }
}
// derivative expressed through the activation OUTPUT y (sign(y) == sign(x) for relu / leaky relu)
__device__ __forceinline__ float act_bwd_from_out(float y, int act) {
    switch (act) {
        case PPOAF_ACT_RELU: return y > 0.f ? 1.f : 0.f;
        case PPOAF_ACT_LEAKY_RELU: return y > 0.f ? 1.f : 0.01f;
        case PPOAF_ACT_TANH: return 1.f - y * y;
        default: return 1.f;
    }
}

enum { EPI_FWD = 0, EPI_BWD_X = 1, EPI_BWD_W = 2 };

struct GemmArgs {
    const float* A; int lda;        // A(m, r): A_RED_CONTIG ? A[rowA(m)*lda + r] : A[rowA(r)*lda + m]
    const float* B; int ldb;        // B(r, n): B_RED_CONTIG ? B[n*ldb + r]       : B[rowB(r)*ldb + n]
    float* C; int ldc;              // C[m*ldc + n]
    int M, N, K;                    // C is M x N, reduction length K
    const int64_t* idxA;            // optional row indirection on A's row index
    const int64_t* idxB;            // optional row indirection on B's row index
    const int32_t* cursor;          // optional device scalar: idx tables start at (*cursor) * cursor_stride
    int cursor_stride;              //   (the minibatch cursor, so one captured graph serves every minibatch)
    const float* bias;              // EPI_FWD: bias[n]
    const float* aux; int ldaux;    // EPI_BWD_X: activation output of the layer below, aux[m*ldaux + n]
    float* dbias;                   // EPI_BWD_W: dbias[m] = sum_r A(m, r)
    int act;
};

// Stage one operand tile (OUT x BK) from global memory into registers.  A "slot" is V consecutive
// floats along the operand's contiguous global dimension: the reduction dim when RED_CONTIG,
// otherwise the output dim.  Out-of-range elements are zero.
template <int OUT, int BK, int NT, bool RED_CONTIG, int V>
__device__ __forceinline__ void stage_tile(float (&regs)[(OUT * BK / V + NT - 1) / NT][V], const float* __restrict__ P,
                                           int ld, const int64_t* __restrict__ idx, int out0, int out_ext, int r0,
                                           int red_ext, int tid) {
    constexpr int SLOTS = (OUT * BK / V + NT - 1) / NT;
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        const int slot = tid + s * NT;
#pragma unroll
        for (int j = 0; j < V; ++j) regs[s][j] = 0.f;
        if (slot >= OUT * BK / V) continue;
        int o, r;
        if constexpr (RED_CONTIG) { o = slot / (BK / V); r = (slot % (BK / V)) * V; }
        else                      { r = slot / (OUT / V); o = (slot % (OUT / V)) * V; }
        const int go = out0 + o, gr = r0 + r;
        if (go >= out_ext || gr >= red_ext) continue;
        const int row_sel = RED_CONTIG ? go : gr;
        const int col_sel = RED_CONTIG ? gr : go;
        const int64_t row = idx ? idx[row_sel] : int64_t(row_sel);
        const float* p = P + row * ld + col_sel;
        if constexpr (V == 4) {
            const float4 t = *reinterpret_cast<const float4*>(p);
            regs[s][0] = t.x; regs[s][1] = t.y; regs[s][2] = t.z; regs[s][3] = t.w;
        } else {
            regs[s][0] = *p;
        }
    }
}

// Commit staged registers into the k-major shared tile S[BK][OUT + 4].
template <int OUT, int BK, int NT, bool RED_CONTIG, int V>
__device__ __forceinline__ void commit_tile(const float (&regs)[(OUT * BK / V + NT - 1) / NT][V],
                                            float (*S)[OUT + 4], int tid) {
    constexpr int SLOTS = (OUT * BK / V + NT - 1) / NT;
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        const int slot = tid + s * NT;
        if (slot >= OUT * BK / V) continue;
        if constexpr (RED_CONTIG) {
            const int o = slot / (BK / V), r = (slot % (BK / V)) * V;
#pragma unroll
            for (int j = 0; j < V; ++j) S[r + j][o] = regs[s][j];
        } else {
            const int r = slot / (OUT / V), o = (slot % (OUT / V)) * V;
            if constexpr (V == 4) {
                *reinterpret_cast<float4*>(&S[r][o]) = make_float4(regs[s][0], regs[s][1], regs[s][2], regs[s][3]);
            } else {
                S[r][o] = regs[s][0];
            }
        }
    }
}

template <int BM, int BN, int BK, bool A_RED_CONTIG, bool B_RED_CONTIG, int VA, int VB, int EPI>
__global__ void __launch_bounds__((BM / 4) * (BN / 4)) gemm_tile_kernel(const GemmArgs g) {
    constexpr int NT = (BM / 4) * (BN / 4);
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];

    const int tid = threadIdx.x;
    const int tx = tid % (BN / 4), ty = tid / (BN / 4);
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

    const int64_t idx_off = g.cursor ? int64_t(*g.cursor) * g.cursor_stride : 0;
    const int64_t* idxA = g.idxA ? g.idxA + idx_off : nullptr;
    const int64_t* idxB = g.idxB ? g.idxB + idx_off : nullptr;

    float ra[(BM * BK / VA + NT - 1) / NT][VA];
    float rb[(BN * BK / VB + NT - 1) / NT][VB];

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float bsum[4] = {0.f, 0.f, 0.f, 0.f};

    const int n_tiles = (g.K + BK - 1) / BK;
    stage_tile<BM, BK, NT, A_RED_CONTIG, VA>(ra, g.A, g.lda, idxA, m0, g.M, 0, g.K, tid);
    stage_tile<BN, BK, NT, B_RED_CONTIG, VB>(rb, g.B, g.ldb, idxB, n0, g.N, 0, g.K, tid);
    commit_tile<BM, BK, NT, A_RED_CONTIG, VA>(ra, As[0], tid);
    commit_tile<BN, BK, NT, B_RED_CONTIG, VB>(rb, Bs[0], tid);
    __syncthreads();
    for (int t = 0; t < n_tiles; ++t) {
        const int buf = t & 1;
        if (t + 1 < n_tiles) {
            stage_tile<BM, BK, NT, A_RED_CONTIG, VA>(ra, g.A, g.lda, idxA, m0, g.M, (t + 1) * BK, g.K, tid);
            stage_tile<BN, BK, NT, B_RED_CONTIG, VB>(rb, g.B, g.ldb, idxB, n0, g.N, (t + 1) * BK, g.K, tid);
        }
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
                if constexpr (EPI == EPI_BWD_W) bsum[i] += av[i];
            }
        }
        if (t + 1 < n_tiles) {
            commit_tile<BM, BK, NT, A_RED_CONTIG, VA>(ra, As[buf ^ 1], tid);
            commit_tile<BN, BK, NT, B_RED_CONTIG, VB>(rb, Bs[buf ^ 1], tid);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= g.M) continue;
        if constexpr (EPI == EPI_BWD_W) {
            if (blockIdx.x == 0 && tx == 0 && g.dbias) g.dbias[m] = bsum[i];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= g.N) continue;
            float v = acc[i][j];
            if constexpr (EPI == EPI_FWD) v = act_fwd(v + g.bias[n], g.act);
            if constexpr (EPI == EPI_BWD_X) v *= act_bwd_from_out(g.aux[int64_t(m) * g.ldaux + n], g.act);
            g.C[int64_t(m) * g.ldc + n] = v;
        }
    }
}

}  // namespace ppoaf
