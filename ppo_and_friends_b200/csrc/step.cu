// One PPO minibatch as a fixed kernel sequence (reference ppo.py:2292-2468 -> policies/ppo_policy.py:
// 891-952, 1012-1055), plus the library-wide utilities (error string, device info).
//
//   grads:  actor forward || critic forward  ->  fused loss fwd/bwd  ->  actor backward || critic backward
//   apply:  [caller all-reduces `grads` when R > 1]  ->  grad sum-of-squares + step scalars  ->  clip + Adam
//
// The actor and critic chains are independent until the loss and again after it, so they run on two
// streams joined by events (capturable: the fork/join becomes graph branches).  Every kernel reads
// the minibatch cursor from device memory, so ONE captured graph serves every full minibatch.
#include <stdarg.h>

#include "internal.h"

namespace ppoaf {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int cached = 0;
    if (cached > 0) return cached;
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) {
        cached = n;
        return n;
    }
    (void)cudaGetLastError();
    return 148;  // B200; only used for sizing queries when no device is visible
}

struct SideStream {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join_fwd = nullptr, fork_bwd = nullptr, join_bwd = nullptr;
    int device = -1;
};

static int get_side_stream(SideStream** out) {
    static thread_local SideStream ss;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    PPOAF_CHECK_ARG(e == cudaSuccess, "no CUDA device: %s", cudaGetErrorString(e));
    if (ss.stream == nullptr || ss.device != dev) {
        e = cudaStreamCreateWithFlags(&ss.stream, cudaStreamNonBlocking);
        PPOAF_CHECK_ARG(e == cudaSuccess, "cudaStreamCreate failed: %s", cudaGetErrorString(e));
        cudaEvent_t* evs[4] = {&ss.fork, &ss.join_fwd, &ss.fork_bwd, &ss.join_bwd};
        for (auto ev : evs) {
            e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
            PPOAF_CHECK_ARG(e == cudaSuccess, "cudaEventCreate failed: %s", cudaGetErrorString(e));
        }
        ss.device = dev;
    }
    *out = &ss;
    return 0;
}

// ---- workspace carving ------------------------------------------------------------------------------
struct NetScratch {
    float* act[PPOAF_MAX_LAYERS + 1];  // act[l] = output of layer l-1 (act[0] unused: the input is gathered)
    float* dz[PPOAF_MAX_LAYERS + 1];   // dz[l]  = gradient w.r.t. the pre-activation of layer l-1's output
};
struct StepScratch {
    NetScratch actor, critic;
    float* loss_partials;
    unsigned int* loss_ticket;
    void* optim_ws;
    size_t total;
};

static void carve(const ppoaf_update_cfg* cfg, int max_batch, char* base, StepScratch* out) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        char* p = base ? base + off : nullptr;
        off += align_up(bytes, 256);
        return p;
    };
    // control blocks first, at offsets that do not depend on the batch: they hold self-resetting tickets
    // that must stay zero between launches, so no activation buffer may ever alias them
    out->loss_ticket = reinterpret_cast<unsigned int*>(take(256));
    out->optim_ws = take(optim_workspace_bytes(0));
    out->loss_partials = reinterpret_cast<float*>(take(loss_workspace_bytes(max_batch, cfg->act_dim)));
    const ppoaf_mlp_desc* nets[2] = {&cfg->actor, &cfg->critic};
    NetScratch* ns[2] = {&out->actor, &out->critic};
    for (int k = 0; k < 2; ++k) {
        for (int l = 1; l <= nets[k]->n_layers; ++l) {
            const size_t rows = (l == nets[k]->n_layers && k == 1) ? size_t(2) * max_batch : size_t(max_batch);
            ns[k]->act[l] = reinterpret_cast<float*>(take(size_t(max_batch) * nets[k]->dims[l] * sizeof(float)));
            ns[k]->dz[l] = reinterpret_cast<float*>(take(rows * nets[k]->dims[l] * sizeof(float)));
        }
    }
    out->total = off;
}

static int check_cfg(const ppoaf_update_cfg* cfg, const char* who) {
    PPOAF_CHECK_ARG(cfg != nullptr, "%s: null cfg", who);
    if (check_mlp_desc(&cfg->actor, who) || check_mlp_desc(&cfg->critic, who)) return 1;
    PPOAF_CHECK_ARG(cfg->critic.dims[cfg->critic.n_layers] == 1, "%s: critic output width must be 1", who);
    PPOAF_CHECK_ARG(cfg->head == PPOAF_HEAD_GAUSSIAN_TANH || cfg->head == PPOAF_HEAD_CATEGORICAL, "%s: unknown head", who);
    if (cfg->head == PPOAF_HEAD_GAUSSIAN_TANH)
        PPOAF_CHECK_ARG(cfg->actor.dims[cfg->actor.n_layers] == cfg->act_dim,
                        "%s: Gaussian actor output width must equal act_dim", who);
    else
        PPOAF_CHECK_ARG(cfg->act_dim == 1, "%s: Categorical actions are stored as one int64 index", who);
    return 0;
}

}  // namespace ppoaf

using namespace ppoaf;

extern "C" int ppoaf_abi_version(void) { return PPOAF_ABI_VERSION; }
extern "C" const char* ppoaf_last_error(void) { return g_err; }

extern "C" int ppoaf_device_info(int* sm, int* cc_major, int* cc_minor) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        set_error("ppoaf_device_info: no CUDA device: %s", cudaGetErrorString(e));
        return 1;
    }
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    PPOAF_CHECK_ARG(e == cudaSuccess, "ppoaf_device_info: %s", cudaGetErrorString(e));
    if (sm) *sm = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return 0;
}

extern "C" int ppoaf_runtime_init(void) {
    SideStream* ss;
    return get_side_stream(&ss);
}

extern "C" size_t ppoaf_update_workspace_bytes(const ppoaf_update_cfg* cfg, int32_t max_batch) {
    if (!cfg || max_batch <= 0) return 0;
    StepScratch s;
    carve(cfg, max_batch, nullptr, &s);
    return s.total;
}

static int run_net_forward(const ppoaf_mlp_desc* net, const float* params, const int64_t* off, const float* x,
                           const ppoaf_update_bufs* b, NetScratch* ns, cudaStream_t s) {
    for (int l = 0; l < net->n_layers; ++l) {
        const bool first = l == 0, last = l + 1 == net->n_layers;
        linear_forward(first ? x : ns->act[l], net->dims[l], first ? b->perm : nullptr, first ? b->mb_cursor : nullptr,
                       b->batch_size, params + off[2 * l], params + off[2 * l + 1], ns->act[l + 1], b->batch,
                       net->dims[l], net->dims[l + 1], last ? PPOAF_ACT_IDENTITY : net->activation, s);
        PPOAF_CHECK_LAUNCH("linear_forward");
    }
    return 0;
}

static int run_net_backward(const ppoaf_mlp_desc* net, const float* params, float* grads, const int64_t* off,
                            const float* x, const ppoaf_update_bufs* b, NetScratch* ns, cudaStream_t s) {
    for (int l = net->n_layers - 1; l >= 0; --l) {
        const bool first = l == 0;
        linear_backward_w(ns->dz[l + 1], first ? x : ns->act[l], net->dims[l], first ? b->perm : nullptr,
                          first ? b->mb_cursor : nullptr, b->batch_size, grads + off[2 * l], grads + off[2 * l + 1],
                          b->batch, net->dims[l], net->dims[l + 1], s);
        PPOAF_CHECK_LAUNCH("linear_backward_w");
        if (!first) {
            linear_backward_x(ns->dz[l + 1], params + off[2 * l], ns->act[l], ns->dz[l], b->batch, net->dims[l],
                              net->dims[l + 1], net->activation, s);
            PPOAF_CHECK_LAUNCH("linear_backward_x");
        }
    }
    return 0;
}

#define PPOAF_CUDA_OK(expr, what)                                                      \
    do {                                                                               \
        cudaError_t e__ = (expr);                                                      \
        if (e__ != cudaSuccess) {                                                      \
            set_error("%s: %s", what, cudaGetErrorString(e__));                        \
            return 2;                                                                  \
        }                                                                              \
    } while (0)

extern "C" int ppoaf_ppo_minibatch_grads(const ppoaf_update_cfg* cfg, const ppoaf_update_bufs* b, void* stream) {
    if (check_cfg(cfg, "ppoaf_ppo_minibatch_grads")) return 1;
    PPOAF_CHECK_ARG(b != nullptr && b->batch >= 1 && b->batch <= b->batch_size && b->n_flat > 0,
                    "ppoaf_ppo_minibatch_grads: bad batch sizes");
    if (b->batch == 1) return 0;  // the reference skips one-row minibatches (ppo.py:2305)
    PPOAF_CHECK_ARG(b->workspace_bytes >= ppoaf_update_workspace_bytes(cfg, b->batch_size),
                    "ppoaf_ppo_minibatch_grads: workspace too small");
    PPOAF_CHECK_ARG(reinterpret_cast<uintptr_t>(b->workspace) % 256 == 0, "ppoaf_ppo_minibatch_grads: workspace must be 256-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    SideStream* ss;
    if (get_side_stream(&ss)) return 1;

    StepScratch sc;
    carve(cfg, b->batch_size, reinterpret_cast<char*>(b->workspace), &sc);  // layout fixed by the nominal B
    const bool gaussian = cfg->head == PPOAF_HEAD_GAUSSIAN_TANH;
    int64_t off_a[2 * PPOAF_MAX_LAYERS + 1], off_c[2 * PPOAF_MAX_LAYERS + 1];
    const int64_t n_actor = param_layout(&cfg->actor, gaussian ? cfg->act_dim : 0, off_a);
    param_layout(&cfg->critic, 0, off_c);
    const float* pa = b->params;
    const float* pc = b->params + n_actor;
    float* ga = b->grads;
    float* gc = b->grads + n_actor;
    const int La = cfg->actor.n_layers, Lc = cfg->critic.n_layers;

    // forward: actor on `s`, critic on the side stream
    PPOAF_CUDA_OK(cudaEventRecord(ss->fork, s), "event record");
    PPOAF_CUDA_OK(cudaStreamWaitEvent(ss->stream, ss->fork, 0), "stream wait");
    if (run_net_forward(&cfg->actor, pa, off_a, b->obs, b, &sc.actor, s)) return 2;
    if (run_net_forward(&cfg->critic, pc, off_c, b->critic_obs, b, &sc.critic, ss->stream)) return 2;
    PPOAF_CUDA_OK(cudaEventRecord(ss->join_fwd, ss->stream), "event record");
    PPOAF_CUDA_OK(cudaStreamWaitEvent(s, ss->join_fwd, 0), "stream wait");

    LossArgs a{};
    a.actor_out = sc.actor.act[La];
    a.critic_out = sc.critic.act[Lc];
    a.log_std = gaussian ? pa + off_a[2 * La] : nullptr;
    a.raw_actions = b->raw_actions;
    a.advantages = b->advantages;
    a.log_probs = b->log_probs;
    a.rewards_to_go = b->rewards_to_go;
    a.values = b->values;
    a.perm = b->perm;
    a.cursor = b->mb_cursor;
    a.batch_size = b->batch_size;
    a.batch = b->batch;
    a.mb_adv_stats = b->mb_adv_stats;
    a.mb_val_stats = b->mb_val_stats;
    a.hparams = b->hparams;
    a.epoch_stats = b->epoch_stats;
    a.d_actor_out = sc.actor.dz[La];
    a.d_critic_out = sc.critic.dz[Lc];
    a.d_log_std = gaussian ? ga + off_a[2 * La] : nullptr;
    a.partials = sc.loss_partials;
    a.ticket = sc.loss_ticket;
    a.head = cfg->head;
    a.act_dim = cfg->act_dim;
    a.pred_dim = cfg->actor.dims[La];
    a.use_huber = cfg->use_huber;
    a.normalize_adv = cfg->normalize_adv;
    a.normalize_values = cfg->normalize_values;
    a.vf_clip_enabled = cfg->vf_clip_enabled;
    a.min_std = cfg->min_std;
    if (launch_ppo_loss(a, s)) return 2;

    // backward: actor on `s`, critic on the side stream
    PPOAF_CUDA_OK(cudaEventRecord(ss->fork_bwd, s), "event record");
    PPOAF_CUDA_OK(cudaStreamWaitEvent(ss->stream, ss->fork_bwd, 0), "stream wait");
    if (run_net_backward(&cfg->actor, pa, ga, off_a, b->obs, b, &sc.actor, s)) return 2;
    if (run_net_backward(&cfg->critic, pc, gc, off_c, b->critic_obs, b, &sc.critic, ss->stream)) return 2;
    PPOAF_CUDA_OK(cudaEventRecord(ss->join_bwd, ss->stream), "event record");
    PPOAF_CUDA_OK(cudaStreamWaitEvent(s, ss->join_bwd, 0), "stream wait");
    return 0;
}

extern "C" int ppoaf_ppo_minibatch_apply(const ppoaf_update_cfg* cfg, const ppoaf_update_bufs* b, void* stream) {
    if (check_cfg(cfg, "ppoaf_ppo_minibatch_apply")) return 1;
    PPOAF_CHECK_ARG(b != nullptr && b->batch >= 1, "ppoaf_ppo_minibatch_apply: bad batch");
    cudaStream_t s = (cudaStream_t)stream;
    if (b->batch == 1) return launch_advance_cursor(b->mb_cursor, s);
    StepScratch sc;
    carve(cfg, b->batch_size, reinterpret_cast<char*>(b->workspace), &sc);
    const bool gaussian = cfg->head == PPOAF_HEAD_GAUSSIAN_TANH;
    const int64_t n_actor = param_layout(&cfg->actor, gaussian ? cfg->act_dim : 0, nullptr);
    const int64_t n_critic = param_layout(&cfg->critic, 0, nullptr);
    return launch_clip_adam(b->params, b->grads, b->adam_m, b->adam_v, b->adam_step, b->mb_cursor, b->hparams, n_actor,
                            n_critic, sc.optim_ws, s);
}
