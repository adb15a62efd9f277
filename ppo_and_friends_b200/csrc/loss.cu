// Fused PPO loss, forward + backward, for one minibatch (reference ppo.py:2299-2438 with the
// distribution arithmetic of networks/distributions.py:491-558,694 and torch's Normal/Categorical).
// One thread per sample; nothing but the gradients w.r.t. the two network outputs, d(log_std) and
// six scalars ever reaches HBM.  Reductions are deterministic: per-CTA partials in fp64, summed in
// CTA order by the last CTA to finish (self-resetting ticket).
//
//   A^   = (A - mean_mb) / (std_mb + 1e-8)                        (mean/std precomputed per epoch)
//   rho  = exp(lp - lp_old);  L_actor = mean(-min(rho A^, clamp(rho, 1-eps, 1+eps) A^)) - w_H mean(H)
//   R^   = (RTG - mu_v) / sqrt(var_v + 1e-8);  L_critic = mean((v - R^)^2) | Huber(delta=10) [| value clip]
//   KL   = mean(lp_old - lp)  (statistic only: as a loss term it is a constant)
#include "internal.h"

namespace ppoaf {

constexpr int kLossThreads = 128;
constexpr int kMaxAct = 64;
enum { LS_ACTOR = 0, LS_CRITIC, LS_CRITIC_CLIPPED, LS_ENTROPY, LS_KL, LS_BAD_RATIO, LS_BAD_VALUE, kLossScalars };

constexpr float kLogSqrt2Pi = 0.91893853320467274178f;

__device__ __forceinline__ float softplus_torch(float x) { return x > 20.f ? x : log1pf(expf(x)); }

__device__ __forceinline__ float critic_term(float v, float target, int use_huber, float& dv) {
    const float d = v - target;
    if (use_huber) {
        const float delta = 10.f, ad = fabsf(d);
        if (ad < delta) { dv = d; return 0.5f * d * d; }
        dv = d > 0.f ? delta : -delta;
        return delta * (ad - 0.5f * delta);
    }
    dv = 2.f * d;
    return d * d;
}

__global__ void __launch_bounds__(kLossThreads) ppo_loss_kernel(const LossArgs a) {
    __shared__ float s_sd[kMaxAct], s_dsd[kMaxAct];
    __shared__ double s_red[kLossThreads / 32][kLossScalars + kMaxAct];
    __shared__ bool s_last;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i = blockIdx.x * kLossThreads + tid;
    const bool live = i < a.batch;
    const int cur = *a.cursor;
    const int64_t* idx = a.perm + int64_t(cur) * a.batch_size;
    const float inv_b = 1.0f / float(a.batch);
    const float w_ent = float(a.hparams[PPOAF_HP_ENTROPY_WEIGHT]);
    const float clip_lo = float(1.0 - a.hparams[PPOAF_HP_SURR_CLIP]), clip_hi = float(1.0 + a.hparams[PPOAF_HP_SURR_CLIP]);
    const float vf_clip = float(a.hparams[PPOAF_HP_VF_CLIP]);
    const bool gaussian = a.head == PPOAF_HEAD_GAUSSIAN_TANH;

    if (gaussian && tid < a.act_dim) {
        // std = max(softplus(log_std), min_std)  (distributions.py:514-515); d std / d log_std
        const float ls = a.log_std[tid];
        const float sp = softplus_torch(ls);
        s_sd[tid] = fmaxf(sp, a.min_std);
        const float sig = ls > 20.f ? 1.f : 1.f / (1.f + expf(-ls));
        s_dsd[tid] = sp > a.min_std ? sig : (sp == a.min_std ? 0.5f * sig : 0.f);
    }
    __syncthreads();

    double sc[kLossScalars];
#pragma unroll
    for (int k = 0; k < kLossScalars; ++k) sc[k] = 0.0;
    float dsd_local[kMaxAct];  // only the first act_dim entries are used (Gaussian)

    if (live) {
        const int64_t j = idx[i];
        float adv = a.advantages[j];
        if (a.normalize_adv) adv = (adv - a.mb_adv_stats[2 * cur]) / a.mb_adv_stats[2 * cur + 1];
        const float lp_old = a.log_probs[j];
        float target = a.rewards_to_go[j];
        if (a.normalize_values) target = (target - a.mb_val_stats[2 * cur]) / a.mb_val_stats[2 * cur + 1];
        const float v = a.critic_out[i];
        a.values[j] = v;                                         // dataset.values[batch_idxs] = values (ppo.py:2340)
        bool bad_value = isnan(v);

        // ---------------- log-prob and entropy ----------------
        float lp = 0.f, ent = 0.f;
        const float* pred = a.actor_out + int64_t(i) * a.pred_dim;
        float pt[kMaxAct], lg[kMaxAct];  // Categorical: renormalised probs and clamped logs
        float cat_S = 1.f;
        int action = 0;
        if (gaussian) {
            const float* x = reinterpret_cast<const float*>(a.raw_actions) + j * a.act_dim;
            float nsum = 0.f, slog = 0.f, ensum = 0.f, eslog = 0.f;
            for (int d = 0; d < a.act_dim; ++d) {
                const float mu = pred[d], sd = s_sd[d], z = x[d] - mu;
                bad_value |= isnan(mu);
                const float lsd = logf(sd);
                const float nl = -(z * z) / (2.f * (sd * sd)) - lsd - kLogSqrt2Pi;
                nsum += fminf(fmaxf(nl, -100.f), 100.f);
                const float th = tanhf(x[d]);
                slog += logf(fmaxf(1.f - th * th, 1e-6f));
                const float nle = -lsd - kLogSqrt2Pi;               // log N(mu; mu, sd)
                ensum += fminf(fmaxf(nle, -100.f), 100.f);
                const float thm = tanhf(mu);
                eslog += logf(fmaxf(1.f - thm * thm, 1e-6f));
            }
            lp = nsum - slog;
            ent = -(ensum - eslog);                                   // entropy = -log_prob(mean) (:694)
        } else {
            const int n = a.pred_dim;
            action = int(reinterpret_cast<const int64_t*>(a.raw_actions)[j * a.act_dim]);
            float mx = pred[0];
            for (int c = 1; c < n; ++c) mx = fmaxf(mx, pred[c]);
            float se = 0.f;
            for (int c = 0; c < n; ++c) { pt[c] = expf(pred[c] - mx); se += pt[c]; bad_value |= isnan(pred[c]); }
            float S = 0.f;
            for (int c = 0; c < n; ++c) { pt[c] = pt[c] / se; S += pt[c]; }   // softmax inside the actor (:1045)
            cat_S = S;
            const float ceps = 1.1920928955078125e-07f;                        // torch.finfo(float32).eps
            float h = 0.f;
            for (int c = 0; c < n; ++c) {
                lg[c] = pt[c];                                                 // keep raw softmax prob for backward
                const float pn = pt[c] / S;                                    // Categorical renormalises probs
                pt[c] = pn;
                const float q = fminf(fmaxf(pn, ceps), 1.f - ceps);
                const float l = logf(q);
                h += pn * l;
                if (c == action) lp = l;
            }
            ent = -h;
        }

        // ---------------- surrogate ----------------
        const float ratio = expf(lp - lp_old);
        const float s1 = ratio * adv;
        const float s2 = fminf(fmaxf(ratio, clip_lo), clip_hi) * adv;
        const bool bad_ratio = isnan(ratio) || isinf(ratio);
        sc[LS_ACTOR] = double(-fminf(s1, s2));
        sc[LS_KL] = double(lp_old - lp);
        sc[LS_ENTROPY] = double(ent);
        sc[LS_BAD_RATIO] = bad_ratio ? 1.0 : 0.0;
        sc[LS_BAD_VALUE] = bad_value ? 1.0 : 0.0;
        // d(-min(s1,s2))/d lp: gradient flows through s1 when s1 <= s2 (ties: both branches carry half
        // and the clamp passes its half exactly when rho is inside the clip range, which a tie implies)
        const float g_lp = (s1 <= s2) ? -adv * ratio * inv_b : 0.f;
        const float g_ent = -w_ent * inv_b;                              // dL/dH_i

        // ---------------- backward through the head ----------------
        float* dpred = a.d_actor_out + int64_t(i) * a.pred_dim;
        if (gaussian) {
            const float* x = reinterpret_cast<const float*>(a.raw_actions) + j * a.act_dim;
            for (int d = 0; d < a.act_dim; ++d) {
                const float mu = pred[d], sd = s_sd[d], z = x[d] - mu;
                const float var = sd * sd;
                const float nl = -(z * z) / (2.f * var) - logf(sd) - kLogSqrt2Pi;
                const float on = (nl >= -100.f && nl <= 100.f) ? 1.f : 0.f;
                const float nle = -logf(sd) - kLogSqrt2Pi;
                const float one = (nle >= -100.f && nle <= 100.f) ? 1.f : 0.f;
                const float thm = tanhf(mu);
                const float tmask = (1.f - thm * thm >= 1e-6f) ? 1.f : 0.f;
                // dlp/dmu = z/var ; dH/dmu = -2 tanh(mu)
                dpred[d] = g_lp * on * (z / var) + g_ent * tmask * (-2.f * thm);
                // dlp/dsd = z^2/sd^3 - 1/sd ; dH/dsd = 1/sd
                dsd_local[d] = g_lp * on * ((z * z) / (var * sd) - 1.f / sd) + g_ent * one * (1.f / sd);
            }
        } else {
            const int n = a.pred_dim;
            const float ceps = 1.1920928955078125e-07f;
            // G_c = dL/dpt_c ; pt = p / S ; p = softmax(z)
            float G[kMaxAct];
            float gdotp = 0.f;
            for (int c = 0; c < n; ++c) {
                const float pn = pt[c];
                const float q = fminf(fmaxf(pn, ceps), 1.f - ceps);
                const float mask = (pn >= ceps && pn <= 1.f - ceps) ? 1.f : 0.f;
                const float l = logf(q);
                // L depends on pt_c through lg_c (log-prob of the action and the p*log p sum) and directly (H)
                const float dlg = (c == action ? g_lp : 0.f) + g_ent * (-pn);
                G[c] = g_ent * (-l) + dlg * mask / q;
                gdotp += G[c] * lg[c];                                       // lg[] holds the raw softmax probs
            }
            float dp[kMaxAct];
            float dpdotp = 0.f;
            for (int c = 0; c < n; ++c) {
                dp[c] = G[c] / cat_S - gdotp / (cat_S * cat_S);
                dpdotp += dp[c] * lg[c];
            }
            for (int c = 0; c < n; ++c) dpred[c] = lg[c] * (dp[c] - dpdotp);
        }

        // ---------------- critic ----------------
        float dv1;
        const float l1 = critic_term(v, target, a.use_huber, dv1);
        sc[LS_CRITIC] = double(l1);
        if (a.vf_clip_enabled) {
            const float vc = fminf(fmaxf(v, -vf_clip), vf_clip);
            float dv2;
            const float l2 = critic_term(vc, target, a.use_huber, dv2);
            sc[LS_CRITIC_CLIPPED] = double(l2);
            const float pass = (v >= -vf_clip && v <= vf_clip) ? 1.f : 0.f;
            a.d_critic_out[i] = dv1 * inv_b;                                 // combined by vf_select_kernel
            a.d_critic_out[a.batch + i] = dv2 * pass * inv_b;
        } else {
            a.d_critic_out[i] = dv1 * inv_b;
        }
    } else if (gaussian) {
        for (int d = 0; d < a.act_dim; ++d) dsd_local[d] = 0.f;
    }

    // ---------------- CTA reduction (fp64, fixed order) ----------------
    const int n_extra = gaussian ? a.act_dim : 0;
#pragma unroll
    for (int k = 0; k < kLossScalars; ++k) {
        const double w = warp_sum(sc[k]);
        if (lane == 0) s_red[warp][k] = w;
    }
    for (int d = 0; d < n_extra; ++d) {
        const double w = warp_sum(double(live ? dsd_local[d] : 0.f));
        if (lane == 0) s_red[warp][kLossScalars + d] = w;
    }
    __syncthreads();
    const int n_vals = kLossScalars + n_extra;
    double* part = reinterpret_cast<double*>(a.partials);
    if (tid < n_vals) {
        double t = 0.0;
        for (int w = 0; w < kLossThreads / 32; ++w) t += s_red[w][tid];
        part[size_t(blockIdx.x) * (kLossScalars + kMaxAct) + tid] = t;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned int done = atomicAdd(a.ticket, 1u);
        s_last = done == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();

    // ---------------- last CTA: totals -> d(log_std), epoch statistics, value-clip weights -----------
    __shared__ double s_tot[kLossScalars + kMaxAct];
    if (tid < n_vals) {
        double t = 0.0;
        for (unsigned b = 0; b < gridDim.x; ++b) t += __ldcg(&part[size_t(b) * (kLossScalars + kMaxAct) + tid]);
        s_tot[tid] = t;
        if (tid >= kLossScalars) a.d_log_std[tid - kLossScalars] = float(t) * s_dsd[tid - kLossScalars];
    }
    __syncthreads();
    if (tid == 0) {
        const double nb = double(a.batch);
        const float actor_mean = float(s_tot[LS_ACTOR] / nb);
        float critic_mean = float(s_tot[LS_CRITIC] / nb);
        if (a.vf_clip_enabled) {
            // critic_loss = max(loss(v), loss(clamp(v)))  (ppo.py:2427-2436; intended semantics);
            // torch.max of two scalars splits the gradient evenly on a tie.
            const float clipped_mean = float(s_tot[LS_CRITIC_CLIPPED] / nb);
            float w1 = 1.f, w2 = 0.f;
            if (clipped_mean > critic_mean) { w1 = 0.f; w2 = 1.f; critic_mean = clipped_mean; }
            else if (clipped_mean == critic_mean) { w1 = 0.5f; w2 = 0.5f; }
            float* wsel = reinterpret_cast<float*>(part + size_t(gridDim.x) * (kLossScalars + kMaxAct));
            wsel[0] = w1; wsel[1] = w2;
        }
        a.epoch_stats[PPOAF_ST_ACTOR_LOSS] += double(actor_mean);
        a.epoch_stats[PPOAF_ST_CRITIC_LOSS] += double(critic_mean);
        if (w_ent != 0.f) a.epoch_stats[PPOAF_ST_ENTROPY] += double(float(s_tot[LS_ENTROPY] / nb));
        a.epoch_stats[PPOAF_ST_KL] += double(float(s_tot[LS_KL] / nb));
        a.epoch_stats[PPOAF_ST_COUNTER] += 1.0;
        a.epoch_stats[PPOAF_ST_BAD_RATIO] += s_tot[LS_BAD_RATIO];
        a.epoch_stats[PPOAF_ST_BAD_VALUE] += s_tot[LS_BAD_VALUE];
        *a.ticket = 0u;  // ready for the next launch
    }
}

__global__ void vf_select_kernel(float* __restrict__ d_critic_out, int batch, const float* __restrict__ wsel) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    d_critic_out[i] = wsel[0] * d_critic_out[i] + wsel[1] * d_critic_out[batch + i];
}

size_t loss_workspace_bytes(int max_batch, int /*act_dim*/) {
    const size_t blocks = size_t((max_batch + kLossThreads - 1) / kLossThreads);
    return align_up(blocks * (kLossScalars + kMaxAct) * sizeof(double) + 16, 256);
}

int launch_ppo_loss(const LossArgs& a, cudaStream_t s) {
    PPOAF_CHECK_ARG(a.act_dim >= 1 && a.act_dim <= kMaxAct && a.pred_dim >= 1 && a.pred_dim <= kMaxAct,
                    "ppo loss: act_dim / prediction width must be in [1, %d]", kMaxAct);
    PPOAF_CHECK_ARG(a.batch >= 2, "ppo loss: minibatches of fewer than 2 rows are skipped by the caller");
    const int blocks = (a.batch + kLossThreads - 1) / kLossThreads;
    ppo_loss_kernel<<<blocks, kLossThreads, 0, s>>>(a);
    PPOAF_CHECK_LAUNCH("ppo_loss_kernel");
    if (a.vf_clip_enabled) {
        const float* wsel = reinterpret_cast<const float*>(reinterpret_cast<const double*>(a.partials) +
                                                           size_t(blocks) * (kLossScalars + kMaxAct));
        vf_select_kernel<<<(a.batch + 255) / 256, 256, 0, s>>>(a.d_critic_out, a.batch, wsel);
        PPOAF_CHECK_LAUNCH("vf_select_kernel");
    }
    return 0;
}

}  // namespace ppoaf

// ---- stand-alone head evaluation (PPOPolicy.evaluate's distribution half, policies/ppo_policy.py:939-950) ----
namespace ppoaf {

__global__ void head_evaluate_kernel(int head, const float* __restrict__ actor_out, int pred_dim,
                                     const float* __restrict__ log_std, float min_std, const void* __restrict__ actions,
                                     int act_dim, int n_rows, float* __restrict__ lp_out, float* __restrict__ ent_out) {
    __shared__ float s_sd[kMaxAct];
    const bool gaussian = head == PPOAF_HEAD_GAUSSIAN_TANH;
    if (gaussian && threadIdx.x < act_dim) s_sd[threadIdx.x] = fmaxf(softplus_torch(log_std[threadIdx.x]), min_std);
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    const float* pred = actor_out + int64_t(i) * pred_dim;
    float lp = 0.f, ent = 0.f;
    if (gaussian) {
        const float* x = reinterpret_cast<const float*>(actions) + int64_t(i) * act_dim;
        float nsum = 0.f, slog = 0.f, ensum = 0.f, eslog = 0.f;
        for (int d = 0; d < act_dim; ++d) {
            const float mu = pred[d], sd = s_sd[d], z = x[d] - mu, lsd = logf(sd);
            nsum += fminf(fmaxf(-(z * z) / (2.f * (sd * sd)) - lsd - kLogSqrt2Pi, -100.f), 100.f);
            const float th = tanhf(x[d]);
            slog += logf(fmaxf(1.f - th * th, 1e-6f));
            ensum += fminf(fmaxf(-lsd - kLogSqrt2Pi, -100.f), 100.f);
            const float thm = tanhf(mu);
            eslog += logf(fmaxf(1.f - thm * thm, 1e-6f));
        }
        lp = nsum - slog;
        ent = -(ensum - eslog);
    } else {
        const int action = int(reinterpret_cast<const int64_t*>(actions)[int64_t(i) * act_dim]);
        float mx = pred[0];
        for (int c = 1; c < pred_dim; ++c) mx = fmaxf(mx, pred[c]);
        float se = 0.f;
        for (int c = 0; c < pred_dim; ++c) se += expf(pred[c] - mx);
        float S = 0.f;
        for (int c = 0; c < pred_dim; ++c) S += expf(pred[c] - mx) / se;
        const float ceps = 1.1920928955078125e-07f;
        float h = 0.f;
        for (int c = 0; c < pred_dim; ++c) {
            const float pn = (expf(pred[c] - mx) / se) / S;
            const float l = logf(fminf(fmaxf(pn, ceps), 1.f - ceps));
            h += pn * l;
            if (c == action) lp = l;
        }
        ent = -h;
    }
    lp_out[i] = lp;
    if (ent_out) ent_out[i] = ent;
}

}  // namespace ppoaf

extern "C" int ppoaf_head_evaluate(int32_t head, const float* actor_out, int32_t pred_dim, const float* log_std,
                                   float min_std, const void* actions, int32_t act_dim, int32_t n_rows,
                                   float* log_prob_out, float* entropy_out, void* stream) {
    using namespace ppoaf;
    PPOAF_CHECK_ARG(head == PPOAF_HEAD_GAUSSIAN_TANH || head == PPOAF_HEAD_CATEGORICAL, "ppoaf_head_evaluate: unknown head");
    PPOAF_CHECK_ARG(act_dim >= 1 && act_dim <= kMaxAct && pred_dim >= 1 && pred_dim <= kMaxAct,
                    "ppoaf_head_evaluate: widths must be in [1, %d]", kMaxAct);
    PPOAF_CHECK_ARG(n_rows >= 0, "ppoaf_head_evaluate: n_rows < 0");
    if (n_rows == 0) return 0;
    head_evaluate_kernel<<<(n_rows + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        head, actor_out, pred_dim, log_std, min_std, actions, act_dim, n_rows, log_prob_out, entropy_out);
    PPOAF_CHECK_LAUNCH("ppoaf_head_evaluate");
    return 0;
}
