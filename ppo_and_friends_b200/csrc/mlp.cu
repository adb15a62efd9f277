// Host-side launchers for the MLP layer kernels (mlp.cuh), the flat parameter layout and the
// forward-only entry point.
#include "mlp.cuh"
#include "internal.h"

namespace ppoaf {

constexpr int kBM = 32, kBN = 64, kBK = 16;

static inline bool vec4_ok(const float* p, int ld, int contig_extent) {
    return (reinterpret_cast<uintptr_t>(p) % 16 == 0) && (ld % 4 == 0) && (contig_extent % 4 == 0);
}

template <bool ARC, bool BRC, int EPI>
static void launch_gemm(const GemmArgs& g, bool va, bool vb, cudaStream_t s) {
    const dim3 grid((g.N + kBN - 1) / kBN, (g.M + kBM - 1) / kBM);
    const int threads = (kBM / 4) * (kBN / 4);
    if (va && vb)
        gemm_tile_kernel<kBM, kBN, kBK, ARC, BRC, 4, 4, EPI><<<grid, threads, 0, s>>>(g);
    else if (va)
        gemm_tile_kernel<kBM, kBN, kBK, ARC, BRC, 4, 1, EPI><<<grid, threads, 0, s>>>(g);
    else if (vb)
        gemm_tile_kernel<kBM, kBN, kBK, ARC, BRC, 1, 4, EPI><<<grid, threads, 0, s>>>(g);
    else
        gemm_tile_kernel<kBM, kBN, kBK, ARC, BRC, 1, 1, EPI><<<grid, threads, 0, s>>>(g);
}

// Y[rows, out] = act(X[idx][rows, in] W^T + b)
void linear_forward(const float* X, int ldx, const int64_t* idx, const int32_t* cursor, int cursor_stride,
                    const float* W, const float* b, float* Y, int rows, int in, int out, int act, cudaStream_t s) {
    GemmArgs g{};
    g.A = X; g.lda = ldx; g.B = W; g.ldb = in; g.C = Y; g.ldc = out;
    g.M = rows; g.N = out; g.K = in;
    g.idxA = idx; g.idxB = nullptr; g.cursor = cursor; g.cursor_stride = cursor_stride;
    g.bias = b; g.act = act;
    launch_gemm<true, true, EPI_FWD>(g, vec4_ok(X, ldx, in), vec4_ok(W, in, in), s);
}

// dX[rows, in] = (dZ[rows, out] W[out, in]) * act'(Xact[rows, in])
void linear_backward_x(const float* dZ, const float* W, const float* Xact, float* dX, int rows, int in, int out,
                       int act, cudaStream_t s) {
    GemmArgs g{};
    g.A = dZ; g.lda = out; g.B = W; g.ldb = in; g.C = dX; g.ldc = in;
    g.M = rows; g.N = in; g.K = out;
    g.aux = Xact; g.ldaux = in; g.act = act;
    launch_gemm<true, false, EPI_BWD_X>(g, vec4_ok(dZ, out, out), vec4_ok(W, in, in), s);
}

// dW[out, in] = dZ[rows, out]^T X[idx][rows, in] ; db[out] = column sums of dZ
void linear_backward_w(const float* dZ, const float* X, int ldx, const int64_t* idx, const int32_t* cursor,
                       int cursor_stride, float* dW, float* db, int rows, int in, int out, cudaStream_t s) {
    GemmArgs g{};
    g.A = dZ; g.lda = out; g.B = X; g.ldb = ldx; g.C = dW; g.ldc = in;
    g.M = out; g.N = in; g.K = rows;
    g.idxA = nullptr; g.idxB = idx; g.cursor = cursor; g.cursor_stride = cursor_stride;
    g.dbias = db;
    launch_gemm<false, false, EPI_BWD_W>(g, vec4_ok(dZ, out, out), vec4_ok(X, ldx, in), s);
}

__global__ void softmax_rows_kernel(float* __restrict__ y, int rows, int n) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    float* p = y + int64_t(r) * n;
    float mx = p[0];
    for (int j = 1; j < n; ++j) mx = fmaxf(mx, p[j]);
    float s = 0.f;
    for (int j = 0; j < n; ++j) { const float e = expf(p[j] - mx); p[j] = e; s += e; }
    for (int j = 0; j < n; ++j) p[j] = p[j] / s;
}

int64_t param_layout(const ppoaf_mlp_desc* net, int32_t log_std_dim, int64_t* offsets) {
    int64_t off = 0;
    for (int l = 0; l < net->n_layers; ++l) {
        const int64_t w = int64_t(net->dims[l]) * net->dims[l + 1];
        if (offsets) offsets[2 * l] = off;
        off += (w + 3) / 4 * 4;
        if (offsets) offsets[2 * l + 1] = off;
        off += (int64_t(net->dims[l + 1]) + 3) / 4 * 4;
    }
    if (offsets) offsets[2 * net->n_layers] = off;
    if (log_std_dim > 0) off += (int64_t(log_std_dim) + 3) / 4 * 4;
    return off;
}

int check_mlp_desc(const ppoaf_mlp_desc* net, const char* who) {
    PPOAF_CHECK_ARG(net != nullptr, "%s: null mlp desc", who);
    PPOAF_CHECK_ARG(net->n_layers >= 1 && net->n_layers <= PPOAF_MAX_LAYERS, "%s: n_layers out of range", who);
    for (int l = 0; l <= net->n_layers; ++l) PPOAF_CHECK_ARG(net->dims[l] > 0, "%s: dims[%d] must be > 0", who, l);
    PPOAF_CHECK_ARG(net->activation >= PPOAF_ACT_IDENTITY && net->activation <= PPOAF_ACT_TANH,
                    "%s: unknown activation", who);
    return 0;
}

}  // namespace ppoaf

using namespace ppoaf;

extern "C" int64_t ppoaf_param_layout(const ppoaf_mlp_desc* net, int32_t log_std_dim, int64_t* offsets) {
    if (check_mlp_desc(net, "ppoaf_param_layout")) return -1;
    return param_layout(net, log_std_dim, offsets);
}

extern "C" size_t ppoaf_mlp_forward_workspace_bytes(const ppoaf_mlp_desc* net, int32_t n_rows) {
    if (!net || n_rows <= 0) return 64;
    int widest = 0;
    for (int l = 1; l <= net->n_layers; ++l) widest = net->dims[l] > widest ? net->dims[l] : widest;
    return 2 * align_up(size_t(n_rows) * widest * sizeof(float), 256) + 256;
}

extern "C" int ppoaf_mlp_forward(const ppoaf_mlp_desc* net, const float* params, const float* x, const int64_t* idx,
                                 int32_t n_rows, int softmax_out, float* y, void* workspace, size_t workspace_bytes,
                                 void* stream) {
    if (check_mlp_desc(net, "ppoaf_mlp_forward")) return 1;
    PPOAF_CHECK_ARG(n_rows >= 0, "ppoaf_mlp_forward: n_rows < 0");
    if (n_rows == 0) return 0;
    PPOAF_CHECK_ARG(workspace_bytes >= ppoaf_mlp_forward_workspace_bytes(net, n_rows),
                    "ppoaf_mlp_forward: workspace too small");
    PPOAF_CHECK_ARG(reinterpret_cast<uintptr_t>(workspace) % 16 == 0, "ppoaf_mlp_forward: workspace alignment");
    cudaStream_t s = (cudaStream_t)stream;
    int64_t off[2 * PPOAF_MAX_LAYERS + 1];
    param_layout(net, 0, off);
    int widest = 0;
    for (int l = 1; l <= net->n_layers; ++l) widest = net->dims[l] > widest ? net->dims[l] : widest;
    const size_t half = align_up(size_t(n_rows) * widest * sizeof(float), 256);
    float* buf[2] = {reinterpret_cast<float*>(workspace),
                     reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + half)};
    const float* in = x;
    const int64_t* in_idx = idx;
    for (int l = 0; l < net->n_layers; ++l) {
        const bool last = l + 1 == net->n_layers;
        float* out = last ? y : buf[l & 1];
        linear_forward(in, net->dims[l], in_idx, nullptr, 0, params + off[2 * l], params + off[2 * l + 1], out, n_rows,
                       net->dims[l], net->dims[l + 1], last ? PPOAF_ACT_IDENTITY : net->activation, s);
        PPOAF_CHECK_LAUNCH("ppoaf_mlp_forward(layer)");
        in = out;
        in_idx = nullptr;
    }
    if (softmax_out) {
        softmax_rows_kernel<<<(n_rows + 127) / 128, 128, 0, s>>>(y, n_rows, net->dims[net->n_layers]);
        PPOAF_CHECK_LAUNCH("ppoaf_mlp_forward(softmax)");
    }
    return 0;
}
