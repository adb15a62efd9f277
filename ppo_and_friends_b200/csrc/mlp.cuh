// fp32 grouped-GEMM core for the actor/critic MLP layers at reference minibatch sizes (B = 128..512,
// hidden 64..512).  One layer of one network is ~30-50 MFLOP: far too small to be a dense
// tensor-core contraction, and a minibatch step is a chain of ~10 dependent phases,
// so the design goal is latency per phase and launches per step, not peak FLOP/s:
//
//   * GROUPED launches: one launch computes every GEMM of a phase (actor and critic forward layer l;
//     or dW and dX of layer l for both networks), so a step is 8 GEMM launches instead of 22 and
//     each launch fills the 148 SMs with 32x64 output tiles;
//   * 256 threads per CTA = 8 warps, each warp a K-group owning 1/8 of every K chunk with an 8x8
//     register microtile (intra-CTA split-K; partial tiles are summed through shared memory).
//     8x8 is the smallest microtile for which the shared-memory pipe (4 cycles per LDS.128) does
//     not cap the FFMA pipe: 16 LDS.128 per 256 FFMA per warp;
//   * operands stream through a 4-stage cp.async ring with BK = 64, in their GLOBAL orientation
//     (so 16-byte cp.async serves all three GEMM flavours, with zero-fill for ragged edges);
//   * fused epilogues (bias + activation; activation derivative; bias gradient + per-tile sum of
//     squares for the gradient-norm clip) and a fused row gather on the operand indexed by sample,
//     so the gathered minibatch never exists in HBM.
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
constexpr int kMaxGroup = 4;

struct GemmProblem {
    const float* A; const float* B; float* C;
    const float* bias;              // EPI_FWD: bias[n]
    const float* aux;               // EPI_BWD_X: activation output of the layer below, aux[m*ldaux + n]
    float* dbias;                   // EPI_BWD_W: dbias[m] = sum_r A(m, r)
    double* sq_out;                 // EPI_BWD_W: sq_out[tile] = sum of squares of this tile of dW (+ db), or null
    const int64_t* idxA;            // optional row indirection on A's row index
    const int64_t* idxB;            // optional row indirection on B's row index
    int lda, ldb, ldc, ldaux;       // A(m,r): A_RC ? A[row(m)*lda + r] : A[row(r)*lda + m];  B likewise with n
    int M, N, K;                    // C is M x N, reduction length K
    int act;
    int flavour;                    // epi * 4 + (VA == 4) * 2 + (VB == 4)
    int tiles_n;                    // number of tiles along N
    int tile_begin;                 // first linear tile id of this problem inside the launch
};
struct GroupedGemmArgs {
    int n_problems;
    int cursor_stride;              // idx tables start at (*cursor) * cursor_stride (the minibatch cursor, so
    const int32_t* cursor;          //   one captured graph serves every minibatch); cursor may be null
    GemmProblem p[kMaxGroup];
};

constexpr int kBM = 32, kBN = 64, kBK = 64, kStages = 4, kThreads = 512, kGroups = 16;
constexpr int kLdRC = kBK + 4;         // [out][BK+4]: 68 = 4*17, consecutive rows land on distinct 16-byte bank groups
constexpr int kLdOCA = kBM + 4;        // A output-contiguous: [BK][32+4]
constexpr int kLdOCB = kBN + 4;        // B output-contiguous: [BK][64+4]
constexpr int kAFloats = (kBM * kLdRC > kBK * kLdOCA ? kBM * kLdRC : kBK * kLdOCA);
constexpr int kBFloats = (kBN * kLdRC > kBK * kLdOCB ? kBN * kLdRC : kBK * kLdOCB);
constexpr int kStageFloats = kAFloats + kBFloats;
constexpr int kRedLd = kBN + 8;       // 72 = 8 (mod 32): the (ty, tx) lanes of a warp hit 32 distinct banks
constexpr int kRedFloats = kGroups * kBM * kRedLd + kGroups * kBM;
constexpr size_t kGemmSmemBytes =
    sizeof(float) * size_t(kStages * kStageFloats > kRedFloats ? kStages * kStageFloats : kRedFloats);

__device__ __forceinline__ void cp_async_16(float* smem_dst, const float* gmem_src, bool valid) {
    const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    const int bytes = valid ? 16 : 0;   // src-size 0: the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_4(float* smem_dst, const float* gmem_src, bool valid) {
    const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    const int bytes = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(s), "l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Per-thread plan for the cp.async's of one operand (OUT x 64 floats per K chunk): everything that does not
// depend on the chunk (shared-memory offsets, gathered row pointers, bounds) is computed once.
//   RC (reduction-contiguous in global): tile[out][r], element P[row(out0+o)*ld + r0 + r]
//   OC (output-contiguous in global):    tile[r][out], element P[row(r0+r)*ld + out0 + o]
template <int OUT, bool RC, int V>
struct ChunkLoader {
    static constexpr int LD_OC = OUT + 4;
    static constexpr int SLOTS = OUT * kBK / V / kThreads;
    static_assert(OUT * kBK / V % kThreads == 0, "slot count");
    const float* src[SLOTS];     // RC: row base + r ; OC: P + o (row added per chunk)
    int dst[SLOTS];              // float offset inside the tile
    int rr[SLOTS];               // reduction index of the slot inside a chunk
    bool ok[SLOTS];              // RC: row in range ; OC: output column in range
    const float* P; const int64_t* idx; int ld;

    __device__ __forceinline__ void plan(const float* __restrict__ P_, int ld_, const int64_t* __restrict__ idx_,
                                         int out0, int out_ext, int tid) {
        P = P_; idx = idx_; ld = ld_;
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            const int slot = tid + s * kThreads;
            int o, r;
            if constexpr (RC) { o = slot / (kBK / V); r = (slot % (kBK / V)) * V; }
            else              { r = slot / (OUT / V); o = (slot % (OUT / V)) * V; }
            rr[s] = r;
            dst[s] = RC ? o * kLdRC + r : r * LD_OC + o;
            const int go = out0 + o;
            ok[s] = go < out_ext;
            if constexpr (RC) {
                const int64_t row = ok[s] ? (idx ? idx[go] : int64_t(go)) : 0;
                src[s] = P + row * ld + r;
            } else {
                src[s] = P + go;
            }
        }
    }
    __device__ __forceinline__ void issue(float* tile, int r0, int red_ext) const {
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            const int gr = r0 + rr[s];
            const bool valid = ok[s] && gr < red_ext;
            const float* p = P;
            if (valid) {
                if constexpr (RC) p = src[s] + r0;
                else p = src[s] + (idx ? idx[gr] : int64_t(gr)) * ld;
            }
            if constexpr (V == 4) cp_async_16(tile + dst[s], p, valid); else cp_async_4(tile + dst[s], p, valid);
        }
    }
};

__device__ __forceinline__ float fast_tanh(float x) {
    // tanh(x) = 1 - 2 / (exp(2x) + 1); ex2/rcp based, absolute error ~2e-7 (the epilogues' hot math)
    const float t = __expf(2.f * x);
    return 1.f - __fdividef(2.f, t + 1.f);
}
__device__ __forceinline__ float act_fwd_fast(float x, int act) {
    switch (act) {
        case PPOAF_ACT_RELU: return x > 0.f ? x : 0.f;
        case PPOAF_ACT_LEAKY_RELU: return x > 0.f ? x : 0.01f * x;
        case PPOAF_ACT_TANH: return fast_tanh(x);
        default: return x;
    }
}

// Position of microtile element i (0..7) of lane-coordinate t inside an operand tile of OUT rows:
//   RC: t + (OUT/8) * i             (consecutive lanes -> consecutive rows: conflict-free LDS.128 along k)
//   OC: two runs of 4 contiguous    (consecutive lanes -> one contiguous 16-byte-per-lane span)
template <int OUT, bool RC>
__device__ __forceinline__ int frag_pos(int t, int i) {
    if constexpr (RC) return t + (OUT / 8) * i;
    else return (i < 4) ? 4 * t + i : OUT / 2 + 4 * t + (i - 4);
}

// f[i][q] = operand(out_i, kk + q), 8 outputs x 4 k's
template <int OUT, bool RC>
__device__ __forceinline__ void load_frag(float (&f)[8][4], const float* tile, int t, int kk) {
    if constexpr (RC) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 v = *reinterpret_cast<const float4*>(tile + (t + (OUT / 8) * i) * kLdRC + kk);
            f[i][0] = v.x; f[i][1] = v.y; f[i][2] = v.z; f[i][3] = v.w;
        }
    } else {
        constexpr int LD_OC = OUT + 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 lo = *reinterpret_cast<const float4*>(tile + (kk + q) * LD_OC + 4 * t);
            const float4 hi = *reinterpret_cast<const float4*>(tile + (kk + q) * LD_OC + OUT / 2 + 4 * t);
            f[0][q] = lo.x; f[1][q] = lo.y; f[2][q] = lo.z; f[3][q] = lo.w;
            f[4][q] = hi.x; f[5][q] = hi.y; f[6][q] = hi.z; f[7][q] = hi.w;
        }
    }
}

#ifdef PPOAF_GEMM_TIMING
__device__ long long g_gemm_stamps[16];
#define PPOAF_STAMP(k) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_gemm_stamps[k] = clock64(); } while (0)
#else
#define PPOAF_STAMP(k) do {} while (0)
#endif

template <bool A_RC, bool B_RC, int VA, int VB, int EPI>
__device__ __forceinline__ void gemm_tile(const GemmProblem& g, int tile, int64_t idx_off, float* smem) {
    const int tid = threadIdx.x;
    PPOAF_STAMP(1);
    const int grp = tid >> 5, lane = tid & 31;
    const int tx = lane & 7, ty = lane >> 3;                  // 8 column-lanes x 4 row-lanes per warp
    const int m0 = (tile / g.tiles_n) * kBM, n0 = (tile % g.tiles_n) * kBN;
    const int64_t* idxA = g.idxA ? g.idxA + idx_off : nullptr;
    const int64_t* idxB = g.idxB ? g.idxB + idx_off : nullptr;

    const int n_chunks = (g.K + kBK - 1) / kBK;
    ChunkLoader<kBM, A_RC, VA> la;
    ChunkLoader<kBN, B_RC, VB> lb;
    la.plan(g.A, g.lda, idxA, m0, g.M, tid);
    lb.plan(g.B, g.ldb, idxB, n0, g.N, tid);
    auto issue = [&](int c) {
        if (c < n_chunks) {
            float* st = smem + (c % kStages) * kStageFloats;
            la.issue(st, c * kBK, g.K);
            lb.issue(st + kAFloats, c * kBK, g.K);
        }
        cp_async_commit();
    };
#pragma unroll 1
    for (int s = 0; s < kStages - 1; ++s) issue(s);
    PPOAF_STAMP(2);

    // epilogue operands (bias / activation of the layer below) are fetched now: their latency hides behind the loop
    const int er = tid / (kBN / 4), ec4 = (tid % (kBN / 4)) * 4;
    float epi[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int n = n0 + ec4 + j;
        if constexpr (EPI == EPI_FWD) { if (n < g.N) epi[j] = g.bias[n]; }
        if constexpr (EPI == EPI_BWD_X) { if (n < g.N && m0 + er < g.M) epi[j] = g.aux[int64_t(m0 + er) * g.ldaux + n]; }
    }

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    float bsum[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) bsum[i] = 0.f;

    for (int c = 0; c < n_chunks; ++c) {
        cp_async_wait<kStages - 2>();
        __syncthreads();                       // chunk c has landed for everyone; chunk c-1's stage is free
        if (c == 0) PPOAF_STAMP(3);
        issue(c + kStages - 1);
        const float* a_tile = smem + (c % kStages) * kStageFloats;
        const float* b_tile = a_tile + kAFloats;
        const int kbase = grp * (kBK / kGroups);
#pragma unroll 1   // one 8x8x4 block (16 LDS.128 + 256 FFMA) is the whole hot loop body
        for (int kb = 0; kb < kBK / kGroups; kb += 4) {
            float a[8][4], b[8][4];
            load_frag<kBM, A_RC>(a, a_tile, ty, kbase + kb);
            load_frag<kBN, B_RC>(b, b_tile, tx, kbase + kb);
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i][q], b[j][q], acc[i][j]);
                    if constexpr (EPI == EPI_BWD_W) bsum[i] += a[i][q];
                }
        }
    }
    PPOAF_STAMP(4);
    cp_async_wait<0>();
    __syncthreads();                           // pipeline memory is reused for the split-K reduction
    PPOAF_STAMP(5);

    float* red = smem;                         // [8 groups][32][65]
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j)
            red[(grp * kBM + frag_pos<kBM, A_RC>(ty, i)) * kRedLd + frag_pos<kBN, B_RC>(tx, j)] = acc[i][j];
    float* red_b = smem + kGroups * kBM * kRedLd;    // [8 groups][32]
    if constexpr (EPI == EPI_BWD_W) {
        if (tx == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) red_b[grp * kBM + frag_pos<kBM, A_RC>(ty, i)] = bsum[i];
        }
    }
    __syncthreads();
    PPOAF_STAMP(6);

    // 2048 outputs / 512 threads: thread handles row r, columns c4 .. c4+3 (one LDS.128 per K-group)
    static_assert(kBM * kBN / kThreads == 4, "epilogue mapping assumes 4 outputs per thread");
    const int r = er, c4 = ec4;
    const int m = m0 + r;
    float sq = 0.f;
    if (m < g.M) {
        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int gq = 0; gq < kGroups; ++gq) {
            const float4 v = *reinterpret_cast<const float4*>(red + (gq * kBM + r) * kRedLd + c4);
            sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
        }
        const float o4[4] = {sum.x, sum.y, sum.z, sum.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + c4 + j;
            if (n >= g.N) continue;
            float o = o4[j];
            if constexpr (EPI == EPI_FWD) o = act_fwd_fast(o + epi[j], g.act);
            if constexpr (EPI == EPI_BWD_X) o *= act_bwd_from_out(epi[j], g.act);
            if constexpr (EPI == EPI_BWD_W) sq = fmaf(o, o, sq);
            g.C[int64_t(m) * g.ldc + n] = o;
        }
        if constexpr (EPI == EPI_BWD_W) {
            if (n0 == 0 && c4 == 0 && g.dbias) {
                float s = 0.f;
#pragma unroll
                for (int gq = 0; gq < kGroups; ++gq) s += red_b[gq * kBM + r];
                g.dbias[m] = s;
                sq = fmaf(s, s, sq);
            }
        }
    }
    PPOAF_STAMP(7);
    if constexpr (EPI == EPI_BWD_W) {
        if (g.sq_out) {                        // per-tile sum of squares -> gradient-norm clip without a norm pass
            __shared__ double s_sq[kThreads / 32];
            const double w = warp_sum(double(sq));
            if (lane == 0) s_sq[grp] = w;
            __syncthreads();
            if (tid == 0) {
                double t = 0.0;
#pragma unroll
                for (int k = 0; k < kThreads / 32; ++k) t += s_sq[k];
                g.sq_out[tile] = t;
            }
        }
    }
}

__global__ void __launch_bounds__(kThreads) grouped_gemm_kernel(const GroupedGemmArgs args) {
    extern __shared__ __align__(16) float smem[];
    PPOAF_STAMP(0);
    int p = 0;
#pragma unroll
    for (int i = 1; i < kMaxGroup; ++i)
        if (i < args.n_problems && int(blockIdx.x) >= args.p[i].tile_begin) p = i;
    const GemmProblem& g = args.p[p];
    const int tile = int(blockIdx.x) - g.tile_begin;
    const int64_t idx_off = args.cursor ? int64_t(*args.cursor) * args.cursor_stride : 0;
    switch (g.flavour) {
        case EPI_FWD * 4 + 3: gemm_tile<true, true, 4, 4, EPI_FWD>(g, tile, idx_off, smem); break;
        case EPI_FWD * 4 + 2: gemm_tile<true, true, 4, 1, EPI_FWD>(g, tile, idx_off, smem); break;
        case EPI_FWD * 4 + 1: gemm_tile<true, true, 1, 4, EPI_FWD>(g, tile, idx_off, smem); break;
        case EPI_FWD * 4 + 0: gemm_tile<true, true, 1, 1, EPI_FWD>(g, tile, idx_off, smem); break;
        case EPI_BWD_X * 4 + 3: gemm_tile<true, false, 4, 4, EPI_BWD_X>(g, tile, idx_off, smem); break;
        case EPI_BWD_X * 4 + 2: gemm_tile<true, false, 4, 1, EPI_BWD_X>(g, tile, idx_off, smem); break;
        case EPI_BWD_X * 4 + 1: gemm_tile<true, false, 1, 4, EPI_BWD_X>(g, tile, idx_off, smem); break;
        case EPI_BWD_X * 4 + 0: gemm_tile<true, false, 1, 1, EPI_BWD_X>(g, tile, idx_off, smem); break;
        case EPI_BWD_W * 4 + 3: gemm_tile<false, false, 4, 4, EPI_BWD_W>(g, tile, idx_off, smem); break;
        case EPI_BWD_W * 4 + 2: gemm_tile<false, false, 4, 1, EPI_BWD_W>(g, tile, idx_off, smem); break;
        case EPI_BWD_W * 4 + 1: gemm_tile<false, false, 1, 4, EPI_BWD_W>(g, tile, idx_off, smem); break;
        default: gemm_tile<false, false, 1, 1, EPI_BWD_W>(g, tile, idx_off, smem); break;
    }
}

}  // namespace ppoaf
