// Gradient averaging scale, per-network clip_grad_norm_ and Adam over the flat [actor | critic]
// buffers (reference policies/ppo_policy.py:1032-1055, utils/mpi_utils.py:89-111; arithmetic of
// torch nn/utils/clip_grad.py and optim/adam.py `_single_tensor_adam`, eps = 1e-5 set by the
// reference).  Two launches: (1) sum of squares per network -> deterministic fp64 CTA partials;
// the last CTA turns them into this step's scalars (clip coefficients, bias corrections) and
// advances the device-side step / minibatch counters; (2) the element-wise update.
// HBM traffic: 16 B read + 12 B written per parameter (+4 B for the norm pass).
#include "internal.h"

namespace ppoaf {

struct StepScalars {
    float coef[2];        // clip coefficient per net (1 when clipping is off)
    float inv_world;
    float neg_step_size;  // -lr / (1 - beta1^t)
    float bc2_sqrt;       // sqrt(1 - beta2^t)
    float w1;             // 1 - beta1
    float beta2;
    float w2;             // 1 - beta2
    float eps;
    float pad[7];
};

constexpr int kOptThreads = 256;

__device__ __forceinline__ double sumsq_range(const float* __restrict__ g, int64_t lo, int64_t hi, float scale) {
    // lo/hi are multiples of 4 floats (layout pads every tensor), g is 16-byte aligned
    float acc = 0.f;
    double tot = 0.0;
    const int64_t v_lo = lo / 4, v_hi = hi / 4;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    int it = 0;
    for (int64_t v = v_lo + int64_t(blockIdx.x) * blockDim.x + threadIdx.x; v < v_hi;
         v += int64_t(gridDim.x) * blockDim.x) {
        const float4 x = g4[v];
        const float a = x.x * scale, b = x.y * scale, c = x.z * scale, d = x.w * scale;
        acc = fmaf(a, a, acc); acc = fmaf(b, b, acc); acc = fmaf(c, c, acc); acc = fmaf(d, d, acc);
        if (++it == 16) { tot += double(acc); acc = 0.f; it = 0; }
    }
    return tot + double(acc);
}

__global__ void __launch_bounds__(kOptThreads)
grad_sumsq_kernel(const float* __restrict__ grads, int64_t n_actor, int64_t n_critic, const double* __restrict__ hp,
                  int64_t* __restrict__ adam_step, int32_t* __restrict__ mb_cursor, double* __restrict__ partials,
                  unsigned int* __restrict__ ticket, StepScalars* __restrict__ out) {
    __shared__ double s_scr[32];
    __shared__ bool s_last;
    const float inv_world = float(hp[PPOAF_HP_INV_WORLD]);
    const double sa = block_sum(sumsq_range(grads, 0, n_actor, inv_world), s_scr);
    const double sc = block_sum(sumsq_range(grads, n_actor, n_actor + n_critic, inv_world), s_scr);
    if (threadIdx.x == 0) {
        partials[2 * blockIdx.x] = sa;
        partials[2 * blockIdx.x + 1] = sc;
        __threadfence();
        s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last || threadIdx.x != 0) return;
    __threadfence();
    double ta = 0.0, tc = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) { ta += __ldcg(&partials[2 * b]); tc += __ldcg(&partials[2 * b + 1]); }
    const float max_norm = float(hp[PPOAF_HP_GRAD_CLIP]);
    StepScalars s;
    if (max_norm >= 0.f) {
        // clip_coef = max_norm / (total_norm + 1e-6), clamped to 1 (torch clip_grad_norm_)
        s.coef[0] = fminf(max_norm / (float(sqrt(ta)) + 1e-6f), 1.f);
        s.coef[1] = fminf(max_norm / (float(sqrt(tc)) + 1e-6f), 1.f);
    } else {
        s.coef[0] = s.coef[1] = 1.f;
    }
    const int64_t t = *adam_step + 1;
    const double b1d = hp[PPOAF_HP_BETA1], b2d = hp[PPOAF_HP_BETA2];
    const double bc1 = 1.0 - pow(b1d, double(t));
    const double bc2 = 1.0 - pow(b2d, double(t));
    s.inv_world = inv_world;
    s.neg_step_size = float(-(hp[PPOAF_HP_LR] / bc1));
    s.bc2_sqrt = float(sqrt(bc2));
    s.w1 = float(1.0 - b1d);
    s.beta2 = float(b2d);
    s.w2 = float(1.0 - b2d);
    s.eps = float(hp[PPOAF_HP_ADAM_EPS]);
    *out = s;
    *adam_step = t;
    if (mb_cursor) *mb_cursor += 1;
    *ticket = 0u;
}

__global__ void __launch_bounds__(kOptThreads)
adam_update_kernel(float* __restrict__ params, const float* __restrict__ grads, float* __restrict__ m,
                   float* __restrict__ v, int64_t n_actor, int64_t n_total, const StepScalars* __restrict__ sp) {
    const StepScalars s = *sp;
    const int64_t nv = n_total / 4, na = n_actor / 4;
    float4* p4 = reinterpret_cast<float4*>(params);
    const float4* g4 = reinterpret_cast<const float4*>(grads);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < nv; i += int64_t(gridDim.x) * blockDim.x) {
        const float coef = s.coef[i < na ? 0 : 1];
        const float4 gq = g4[i];
        float4 pq = p4[i], mq = m4[i], vq = v4[i];
        float g[4] = {gq.x, gq.y, gq.z, gq.w}, p[4] = {pq.x, pq.y, pq.z, pq.w};
        float mm[4] = {mq.x, mq.y, mq.z, mq.w}, vv[4] = {vq.x, vq.y, vq.z, vq.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            // explicit _rn intrinsics: same operation order as torch's CPU kernels, no FMA contraction
            const float gk = __fmul_rn(__fmul_rn(g[k], s.inv_world), coef);
            mm[k] = __fadd_rn(mm[k], __fmul_rn(s.w1, __fsub_rn(gk, mm[k])));               // lerp_(g, 1-b1)
            vv[k] = __fadd_rn(__fmul_rn(vv[k], s.beta2), __fmul_rn(__fmul_rn(s.w2, gk), gk));  // mul_ ; addcmul_
            const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vv[k]), s.bc2_sqrt), s.eps);
            p[k] = __fadd_rn(p[k], __fdiv_rn(__fmul_rn(s.neg_step_size, mm[k]), denom));    // addcdiv_
        }
        p4[i] = make_float4(p[0], p[1], p[2], p[3]);
        m4[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
        v4[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    }
}

__global__ void advance_cursor_kernel(int32_t* mb_cursor) { *mb_cursor += 1; }

static int opt_grid(int64_t n_total) {
    int64_t b = ceil_div64(n_total / 4, kOptThreads);
    const int64_t cap = int64_t(sm_count()) * 2;
    return int(b < 1 ? 1 : (b > cap ? cap : b));
}

size_t optim_workspace_bytes(int64_t /*n_total*/) {
    return align_up(size_t(sm_count()) * 2 * 2 * sizeof(double), 256) + 256;  // partials | StepScalars | ticket
}

int launch_clip_adam(float* params, const float* grads, float* m, float* v, int64_t* adam_step, int32_t* mb_cursor,
                     const double* hparams, int64_t n_actor, int64_t n_critic, void* workspace, cudaStream_t s) {
    PPOAF_CHECK_ARG(n_actor % 4 == 0 && n_critic % 4 == 0, "clip_adam: segment sizes must be multiples of 4 floats");
    PPOAF_CHECK_ARG((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) |
                     reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) % 16 == 0,
                    "clip_adam: buffers must be 16-byte aligned");
    const int64_t n_total = n_actor + n_critic;
    const int grid = opt_grid(n_total);
    double* partials = reinterpret_cast<double*>(workspace);
    char* tail = reinterpret_cast<char*>(workspace) + align_up(size_t(sm_count()) * 2 * 2 * sizeof(double), 256);
    StepScalars* sc = reinterpret_cast<StepScalars*>(tail);
    unsigned int* ticket = reinterpret_cast<unsigned int*>(tail + 128);
    grad_sumsq_kernel<<<grid, kOptThreads, 0, s>>>(grads, n_actor, n_critic, hparams, adam_step, mb_cursor, partials,
                                                   ticket, sc);
    PPOAF_CHECK_LAUNCH("grad_sumsq_kernel");
    adam_update_kernel<<<grid, kOptThreads, 0, s>>>(params, grads, m, v, n_actor, n_total, sc);
    PPOAF_CHECK_LAUNCH("adam_update_kernel");
    return 0;
}

int launch_advance_cursor(int32_t* mb_cursor, cudaStream_t s) {
    advance_cursor_kernel<<<1, 1, 0, s>>>(mb_cursor);
    PPOAF_CHECK_LAUNCH("advance_cursor_kernel");
    return 0;
}

}  // namespace ppoaf

using namespace ppoaf;

extern "C" int ppoaf_clip_adam_step(float* params, const float* grads, float* adam_m, float* adam_v,
                                    int64_t* adam_step, const double* hparams, int64_t n_actor, int64_t n_critic,
                                    void* workspace, size_t workspace_bytes, void* stream) {
    PPOAF_CHECK_ARG(workspace_bytes >= optim_workspace_bytes(n_actor + n_critic), "ppoaf_clip_adam_step: workspace too small");
    PPOAF_CHECK_ARG(reinterpret_cast<uintptr_t>(workspace) % 16 == 0, "ppoaf_clip_adam_step: workspace alignment");
    return launch_clip_adam(params, grads, adam_m, adam_v, adam_step, nullptr, hparams, n_actor, n_critic, workspace,
                            (cudaStream_t)stream);
}
