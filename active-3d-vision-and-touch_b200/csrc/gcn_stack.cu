// A whole GCN.forward (pterotactyl/reconstruction/vision/model.py:316-331) and its backward as ONE C-ABI call each.
//
// The reference runs 20 x (matmul, dense adjacency matmul, cat, bias add, relu) from Python; the per-layer kernels of
// this library are 5-135 us at the training batch, so a host that issues them one ctypes call at a time cannot keep
// the GPU fed (measured: 19.95 ms per reconstruction step eagerly vs 18.5 ms replayed from a CUDA graph).  These two
// entry points walk the layers in native code: same kernels, same order, same results as calling the per-layer
// entry points (tests/test_gcn_gpu.py checks bit equality), no Python between launches.
#include "ptk_common.cuh"

using namespace ptk;

namespace {

inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

inline bool aligned16(const void *p) { return ((uintptr_t)p % 16) == 0; }

// ops._aggregate: the form with the hub rows' common neighbour set needs the vector path; otherwise the plain CSR
int aggregate(const ptk_gcn_csr *g, int64_t Nv, const float *in, int64_t B, int64_t C, int64_t L, const float *bias,
              int relu, float *out, int64_t ldi, int64_t ldo, ptk_stream_t stream) {
    const bool vector_path = (C % 4 == 0) && L >= 1 && L <= 384 && aligned16(in) && aligned16(out) && aligned16(bias);
    if (g->n_common > 0 && vector_path)
        return ptk_gcn_aggregate_ex(g->k_rowptr, g->k_col, g->k_val, g->k_hubs, g->k_n_hubs, g->common_col, g->common_w,
                                    g->n_common, g->alpha, g->row_skip, Nv, in, B, C, L, bias, relu, out, ldi, ldo,
                                    stream);
    return ptk_gcn_aggregate_ex(g->rowptr, g->col, g->val, g->hubs, g->n_hubs, nullptr, nullptr, 0, nullptr, nullptr, Nv,
                                in, B, C, L, bias, relu, out, ldi, ldo, stream);
}

// ops._fused_layer_ok
inline bool fused_layer_ok(int64_t K, int64_t N, int64_t Lc, int relu) {
    const int64_t Lp = (Lc + 3) / 4 * 4;
    return relu && K % 4 == 0 && N % 4 == 0 && N >= 64 && Lc >= 1 && Lp < N && Lp <= 384;
}

struct FwdPlan {
    size_t head_off, h_off, lin_off, total;
};

FwdPlan plan_fwd(int64_t M, int32_t n, const int64_t *widths, const int32_t *Ls) {
    int64_t max_lp = 4, max_n = 1;
    size_t lin = 0;
    for (int l = 0; l < n; ++l) {
        const int64_t Lp = (Ls[l] + 3) / 4 * 4;
        if (Lp > max_lp) max_lp = Lp;
        if (widths[l + 1] > max_n) max_n = widths[l + 1];
        const size_t b = ptk_gcn_linear_workspace_bytes(M, widths[l], widths[l + 1]);
        if (b > lin) lin = b;
    }
    FwdPlan p;
    p.head_off = 0;
    p.h_off = align_up(sizeof(float) * (size_t)M * (size_t)max_lp);
    p.lin_off = p.h_off + align_up(sizeof(float) * (size_t)M * (size_t)max_n);
    p.total = p.lin_off + align_up(lin) + 256;
    return p;
}

struct BwdPlan {
    size_t gh_off, tmp_off, lin_off, bg_off, slab_off, total;
    size_t lin_bytes, bg_bytes;
    int group_n;          // layers whose output gradients live in the slab (0: none)
    int64_t group_w;      // their width
    int32_t group_L;      // their propagated width
};

// The layers 0..n-2 that share (output width, propagated width) with layer n-2 write the gradient of their output into
// one slab, so that their bias gradients are two batched launches at the end instead of two per layer (ops._GCNStack).
BwdPlan plan_bwd(int64_t M, int32_t n, const int64_t *widths, const int32_t *Ls, const uint8_t *need_gb, int batch_bias) {
    BwdPlan p;
    int64_t max_w = 1;
    size_t lin = 0, bg = 0;
    for (int l = 0; l <= n; ++l)
        if (widths[l] > max_w) max_w = widths[l];
    for (int l = 0; l < n; ++l) {
        size_t b = ptk_gcn_linear_workspace_bytes(M, widths[l], widths[l + 1]);
        const size_t w = ptk_gcn_linear_wgrad_workspace_bytes(M, widths[l], widths[l + 1]);
        if (w > b) b = w;
        if (b > lin) lin = b;
        const size_t g = ptk_gcn_bias_grad_workspace_bytes(M, Ls[l]);
        if (g > bg) bg = g;
    }
    p.group_n = 0;
    p.group_w = 0;
    p.group_L = 0;
    if (batch_bias && n >= 3) {
        const int64_t rw = widths[n - 1];
        const int32_t rl = Ls[n - 2];
        int cnt = 0;
        for (int l = 0; l < n - 1; ++l)
            if (need_gb[l] && widths[l + 1] == rw && Ls[l] == rl) ++cnt;
        if (cnt >= 2) {
            p.group_n = cnt;
            p.group_w = rw;
            p.group_L = rl;
            const size_t b = ptk_gcn_bias_grad_batched_workspace_bytes(cnt, M, rl);
            if (b > bg) bg = b;
        }
    }
    p.lin_bytes = align_up(lin);
    p.bg_bytes = align_up(bg);
    p.gh_off = 0;
    p.tmp_off = align_up(sizeof(float) * (size_t)M * (size_t)max_w);
    p.lin_off = p.tmp_off + align_up(sizeof(float) * (size_t)M * (size_t)max_w);
    p.bg_off = p.lin_off + p.lin_bytes;
    p.slab_off = p.bg_off + p.bg_bytes;
    p.total = p.slab_off + align_up(sizeof(float) * (size_t)p.group_n * (size_t)M * (size_t)p.group_w) + 256;
    return p;
}

inline bool in_group(const BwdPlan &p, int l, int n, const int64_t *widths, const int32_t *Ls, const uint8_t *need_gb) {
    return p.group_n > 0 && l >= 0 && l < n - 1 && need_gb[l] && widths[l + 1] == p.group_w && Ls[l] == p.group_L;
}

}  // namespace

extern "C" size_t ptk_gcn_stack_fwd_workspace_bytes(int64_t B, int64_t Nv, int32_t n_layers, const int64_t *widths,
                                                    const int32_t *Ls) {
    if (B <= 0 || Nv <= 0 || n_layers <= 0 || !widths || !Ls) return 0;
    return plan_fwd(B * Nv, n_layers, widths, Ls).total;
}

extern "C" int ptk_gcn_stack_fwd(const ptk_gcn_csr *graph, int64_t B, int64_t Nv, int32_t n_layers,
                                 const int64_t *widths, const int32_t *Ls, const uint8_t *relus, const float *X,
                                 const float *const *W, const float *const *bias, float *const *acts,
                                 uint32_t *const *x_bits, int algo, int fuse, void *workspace, size_t workspace_bytes,
                                 ptk_stream_t stream) {
    PTK_REQUIRE(graph && widths && Ls && relus && X && W && bias && acts, PTK_ERR_SHAPE, "gcn_stack_fwd: null pointer");
    PTK_REQUIRE(B > 0 && Nv > 0 && n_layers > 0 && n_layers <= 4096, PTK_ERR_SHAPE, "gcn_stack_fwd: bad sizes");
    const int64_t M = B * Nv;
    const FwdPlan p = plan_fwd(M, n_layers, widths, Ls);
    PTK_REQUIRE(workspace && workspace_bytes >= p.total, PTK_ERR_WORKSPACE, "gcn_stack_fwd: workspace too small");
    PTK_REQUIRE(aligned16(workspace), PTK_ERR_ALIGN, "gcn_stack_fwd: workspace must be 16-byte aligned");
    char *ws = reinterpret_cast<char *>(workspace);
    float *head = reinterpret_cast<float *>(ws + p.head_off);
    float *Hbuf = reinterpret_cast<float *>(ws + p.h_off);
    void *lin = ws + p.lin_off;
    const size_t lin_bytes = p.total - p.lin_off;
    const float *in = X;
    for (int l = 0; l < n_layers; ++l) {
        const int64_t K = widths[l], N = widths[l + 1];
        const int32_t Lc = Ls[l];
        PTK_REQUIRE(W[l] && bias[l] && acts[l] && K > 0 && N > 0 && Lc >= 0 && Lc <= N, PTK_ERR_SHAPE,
                    "gcn_stack_fwd: layer %d: bad arguments", l);
        int rc;
        if (fuse && algo == PTK_GEMM_FFMA && fused_layer_ok(K, N, Lc, relus[l]) && aligned16(bias[l]) &&
            aligned16(in) && aligned16(W[l]) && aligned16(acts[l])) {
            // one GCN_layer.forward in two kernels: the exact GEMM writes H[:, :Lp] into the compact head and
            // relu(H[:, Lp:]) into the output (and packs the ReLU mask of its INPUT for this layer's dgrad); the
            // aggregation gathers from the head and fills out[:, :, :Lp]
            const int64_t Lp = (Lc + 3) / 4 * 4;
            uint32_t *xb = (x_bits && K <= 512) ? x_bits[l] : nullptr;
            rc = ptk_gcn_linear_fwd_split(in, W[l], M, K, N, Lp, head, acts[l], 1, xb, stream);
            if (rc) return rc;
            rc = aggregate(graph, Nv, head, B, Lp, Lc, bias[l], 1, acts[l], Lp, N, stream);
            if (rc) return rc;
        } else {
            rc = ptk_gcn_linear_fwd(in, W[l], M, K, N, Hbuf, algo, lin, lin_bytes, stream);
            if (rc) return rc;
            rc = aggregate(graph, Nv, Hbuf, B, N, Lc, bias[l], relus[l], acts[l], 0, 0, stream);
            if (rc) return rc;
        }
        in = acts[l];
    }
    return PTK_OK;
}

extern "C" size_t ptk_gcn_stack_bwd_workspace_bytes(int64_t B, int64_t Nv, int32_t n_layers, const int64_t *widths,
                                                    const int32_t *Ls, const uint8_t *need_gb, int batch_bias) {
    if (B <= 0 || Nv <= 0 || n_layers <= 0 || !widths || !Ls || !need_gb) return 0;
    return plan_bwd(B * Nv, n_layers, widths, Ls, need_gb, batch_bias).total;
}

extern "C" int ptk_gcn_stack_bwd(const ptk_gcn_csr *graph_t, int64_t B, int64_t Nv, int32_t n_layers,
                                 const int64_t *widths, const int32_t *Ls, const uint8_t *relus, const float *X,
                                 const float *const *W, const float *const *acts, const uint32_t *const *x_bits,
                                 const float *gout, float *gX, float *const *gW, float *const *gb,
                                 const uint8_t *need_gb, int batch_bias, int algo_dgrad, int algo_wgrad,
                                 void *workspace, size_t workspace_bytes, ptk_stream_t stream) {
    PTK_REQUIRE(graph_t && widths && Ls && relus && X && W && acts && gout && gW && gb && need_gb, PTK_ERR_SHAPE,
                "gcn_stack_bwd: null pointer");
    PTK_REQUIRE(B > 0 && Nv > 0 && n_layers > 0 && n_layers <= 4096, PTK_ERR_SHAPE, "gcn_stack_bwd: bad sizes");
    const int n = n_layers;
    const int64_t M = B * Nv;
    BwdPlan p = plan_bwd(M, n, widths, Ls, need_gb, batch_bias);
    PTK_REQUIRE(workspace && workspace_bytes >= p.total, PTK_ERR_WORKSPACE, "gcn_stack_bwd: workspace too small");
    PTK_REQUIRE(aligned16(workspace), PTK_ERR_ALIGN, "gcn_stack_bwd: workspace must be 16-byte aligned");
    char *ws = reinterpret_cast<char *>(workspace);
    float *gH = reinterpret_cast<float *>(ws + p.gh_off);
    float *tmp = reinterpret_cast<float *>(ws + p.tmp_off);
    void *lin = ws + p.lin_off;
    void *bg = ws + p.bg_off;
    float *slab = reinterpret_cast<float *>(ws + p.slab_off);
    // the batched bias gradients are written as one (group_n, width) matrix: the group's gb pointers must be that
    int first = -1, slot_of[4096];
    if (p.group_n > 0) {
        int s = 0;
        for (int l = 0; l < n - 1; ++l) {
            slot_of[l] = -1;
            if (!in_group(p, l, n, widths, Ls, need_gb)) continue;
            if (first < 0) first = l;
            if (!gb[l] || gb[l] != gb[first] + (size_t)s * p.group_w) {  // not contiguous: per-layer bias gradients
                p.group_n = 0;
                break;
            }
            slot_of[l] = s++;
        }
    }
    auto slot = [&](int l) { return (p.group_n > 0 && l >= 0 && l < n - 1) ? slot_of[l] : -1; };

    const float *g = gout;  // gradient w.r.t. the output of layer l
    int rc;
    if (relus[n - 1]) {
        rc = ptk_relu_mask(gout, acts[n - 1], M * widths[n], tmp, stream);
        if (rc) return rc;
        g = tmp;
    }
    for (int l = n - 1; l >= 0; --l) {
        const int64_t K = widths[l], N = widths[l + 1];
        const float *act_in = l == 0 ? X : acts[l - 1];  // input of layer l
        if (need_gb[l] && slot(l) < 0) {
            PTK_REQUIRE(gb[l], PTK_ERR_SHAPE, "gcn_stack_bwd: layer %d: bias gradient requested without a buffer", l);
            rc = ptk_gcn_bias_grad(g, M, N, Ls[l], gb[l], bg, p.bg_bytes, stream);
            if (rc) return rc;
        }
        rc = aggregate(graph_t, Nv, g, B, N, Ls[l], nullptr, 0, gH, 0, 0, stream);
        if (rc) return rc;
        if (gW[l]) {
            rc = ptk_gcn_linear_wgrad(act_in, gH, M, K, N, gW[l], algo_wgrad, lin, p.lin_bytes, stream);
            if (rc) return rc;
        }
        if (l > 0 || gX) {
            const bool masked = l > 0 && relus[l - 1];
            float *dst = l == 0 ? gX : (slot(l - 1) >= 0 ? slab + (size_t)slot(l - 1) * M * p.group_w : tmp);
            rc = ptk_gcn_linear_dgrad(gH, W[l], masked ? act_in : nullptr,
                                      (masked && x_bits) ? x_bits[l] : nullptr, M, K, N, dst, algo_dgrad, lin,
                                      p.lin_bytes, stream);
            if (rc) return rc;
            g = dst;
        }
    }
    if (p.group_n > 0) {
        rc = ptk_gcn_bias_grad_batched(slab, p.group_n, M, p.group_w, p.group_L, gb[first], bg, p.bg_bytes, stream);
        if (rc) return rc;
    }
    return PTK_OK;
}
