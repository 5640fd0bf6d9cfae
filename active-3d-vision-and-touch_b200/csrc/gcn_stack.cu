// A whole GCN.forward (pterotactyl/reconstruction/vision/model.py:316-331) and its backward as ONE C-ABI call each.
//
// The reference runs 20 x (matmul, dense adjacency matmul, cat, bias add, relu) from Python; the per-layer kernels of
// this library are 5-135 us at the training batch, so a host that issues them one ctypes call at a time cannot keep
// the GPU fed (measured: 19.95 ms per reconstruction step eagerly vs 18.5 ms replayed from a CUDA graph).  These two
// entry points walk the layers in native code: same kernels, same order, same results as calling the per-layer
// entry points (tests/test_gcn_gpu.py checks bit equality), no Python between launches.
#include "ptk_common.cuh"

namespace ptk {
// tensor-core GEMMs (gemm_tf32x3.cu)
bool tf32x3_eligible(const void *A, const void *D, int64_t M, int64_t K, int64_t N);
int gemm_tf32x3_presplit(const float *A, const float *b_hi, const float *b_lo, const float *act, const uint32_t *act_bits,
                         int64_t M, int64_t K, int64_t N, float *D, uint32_t *mask_ws, float *part, int *flags,
                         cudaStream_t st);
size_t tf32x3_part_floats();
int tf32x3_max_ctas();
int tf32x3_presplit_batched(int n, const float *const *src, const int *rows, const int *cols, int transpose,
                            float *const *hi, float *const *lo, int *flags, int n_flags, cudaStream_t st);
bool wgrad_tf32x3_eligible(const void *X, const void *gH, int64_t M, int64_t Kin, int64_t Nout);
size_t wgrad_tf32x3_workspace_bytes(int64_t M, int64_t Kin, int64_t Nout);
int wgrad_tf32x3(const float *X, const float *gH, int64_t M, int64_t Kin, int64_t Nout, void *workspace,
                 size_t workspace_bytes, float **part_out, int *n_splits, cudaStream_t st);

// second stage of the split-M weight gradients of a whole pass in one launch: out[l][e] = sum_s part[l][s, e] in
// ascending s -- the summation order of splitk_reduce_kernel (gcn_linear.cu), so the results are bit-identical
constexpr int RB_MAX = 32;
struct ReduceBatch {
    const float *part[RB_MAX];
    float *out[RB_MAX];
    int ns[RB_MAX];
    long long elems[RB_MAX];
};
__global__ void splitk_reduce_batched_kernel(const __grid_constant__ ReduceBatch b) {
    pdl_launch_dependents();
    pdl_wait();
    const int l = blockIdx.y;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long elems = b.elems[l];
    if (e >= elems) return;
    const float *p = b.part[l];
    float acc = 0.f;
    for (int s = 0; s < b.ns[l]; ++s) acc += p[(size_t)s * elems + e];
    b.out[l][e] = acc;
}
}  // namespace ptk

using namespace ptk;

namespace {

constexpr int MAX_LAYERS = 4096;

inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

inline bool aligned16(const void *p) { return ((uintptr_t)p % 16) == 0; }

// ops._aggregate: the form with the hub rows' common neighbour set needs the vector path; otherwise the plain CSR
int aggregate(const ptk_gcn_csr *g, int64_t Nv, const float *in, int64_t B, int64_t C, int64_t L, const float *bias,
              int relu, float *out, int64_t ldi, int64_t ldo, ptk_stream_t stream) {
    const bool vector_path = (C % 4 == 0) && L >= 1 && L <= 384 && aligned16(in) && aligned16(out) && aligned16(bias);
    if (vector_path && (g->n_common > 0 || g->tile_uptr))
        // the kernel form (== the plain arrays when no common set was split off), rows staged in shared memory per tile
        return ptk_gcn_aggregate_tiled(g->k_rowptr, g->k_col, g->k_val, g->k_hubs, g->k_n_hubs, g->common_col, g->common_w,
                                       g->n_common, g->alpha, g->row_skip, g->tile_uptr, g->tile_ucol, g->tile_lidx,
                                       g->max_union, PTK_AGG_AUTO, Nv, in, B, C, L, bias, relu, out, ldi, ldo, stream);
    return ptk_gcn_aggregate_ex(g->rowptr, g->col, g->val, g->hubs, g->n_hubs, nullptr, nullptr, 0, nullptr, nullptr, Nv,
                                in, B, C, L, bias, relu, out, ldi, ldo, stream);
}

// ops._fused_layer_ok
inline bool fused_layer_ok(int64_t K, int64_t N, int64_t Lc, int relu) {
    const int64_t Lp = (Lc + 3) / 4 * 4;
    return relu && K % 4 == 0 && N % 4 == 0 && N >= 64 && Lc >= 1 && Lp < N && Lp <= 384;
}

struct FwdPlan {
    size_t head_off, h_off, lin_off, split_off, flags_off, part_off, total;
    size_t lin_bytes;
};

// bytes of the pre-split (hi, lo) weights of every layer, laid out back to back
size_t split_bytes(int32_t n, const int64_t *widths) {
    size_t b = 0;
    for (int l = 0; l < n; ++l) b += align_up(2 * sizeof(float) * (size_t)widths[l] * (size_t)widths[l + 1]);
    return b;
}

FwdPlan plan_fwd(int64_t M, int32_t n, const int64_t *widths, const int32_t *Ls, int algo) {
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
    p.lin_bytes = align_up(lin);
    p.split_off = p.lin_off + p.lin_bytes;
    const bool tc = algo != PTK_GEMM_FFMA;  // tensor-core layers take their weights pre-split, in one launch
    p.flags_off = p.split_off + (tc ? split_bytes(n, widths) : 0);
    p.part_off = p.flags_off + (tc ? align_up(sizeof(int) * (size_t)n * (size_t)tf32x3_max_ctas()) : 0);
    p.total = p.part_off + (tc ? align_up(sizeof(float) * tf32x3_part_floats()) : 0) + 256;
    return p;
}

struct BwdPlan {
    size_t gh_off, tmp_off, lin_off, bg_off, slab_off, split_off, flags_off, part_off, mask_off, wg_off, total;
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
    // tensor-core dgrad: every layer's weights pre-split in one launch, own partial-tile flags per layer, one
    // partial-tile buffer, one mask area; tensor-core wgrad: every layer keeps its split-M partial sums until the
    // single batched reduction at the end of the pass
    p.split_off = p.slab_off + align_up(sizeof(float) * (size_t)p.group_n * (size_t)M * (size_t)p.group_w);
    p.flags_off = p.split_off + split_bytes(n, widths);
    p.part_off = p.flags_off + align_up(sizeof(int) * (size_t)n * (size_t)tf32x3_max_ctas());
    p.mask_off = p.part_off + align_up(sizeof(float) * tf32x3_part_floats());
    p.wg_off = p.mask_off + align_up(sizeof(uint32_t) * (size_t)M * (size_t)((max_w + 31) / 32));
    size_t wg = 0;
    for (int l = 0; l < n; ++l) wg += align_up(wgrad_tf32x3_workspace_bytes(M, widths[l], widths[l + 1]));
    p.total = p.wg_off + wg + 256;
    return p;
}

inline bool in_group(const BwdPlan &p, int l, int n, const int64_t *widths, const int32_t *Ls, const uint8_t *need_gb) {
    return p.group_n > 0 && l >= 0 && l < n - 1 && need_gb[l] && widths[l + 1] == p.group_w && Ls[l] == p.group_L;
}

}  // namespace

extern "C" size_t ptk_gcn_stack_fwd_workspace_bytes(int64_t B, int64_t Nv, int32_t n_layers, const int64_t *widths,
                                                    const int32_t *Ls) {
    if (B <= 0 || Nv <= 0 || n_layers <= 0 || !widths || !Ls) return 0;
    return plan_fwd(B * Nv, n_layers, widths, Ls, PTK_GEMM_AUTO).total;  // the larger of the two layouts
}

extern "C" int ptk_gcn_stack_fwd(const ptk_gcn_csr *graph, int64_t B, int64_t Nv, int32_t n_layers,
                                 const int64_t *widths, const int32_t *Ls, const uint8_t *relus, const float *X,
                                 const float *const *W, const float *const *bias, float *const *acts,
                                 uint32_t *const *x_bits, int algo, int fuse, void *workspace, size_t workspace_bytes,
                                 ptk_stream_t stream) {
    PTK_NVTX("ptk_gcn_stack_fwd");
    PTK_REQUIRE(graph && widths && Ls && relus && X && W && bias && acts, PTK_ERR_SHAPE, "gcn_stack_fwd: null pointer");
    PTK_REQUIRE(B > 0 && Nv > 0 && n_layers > 0 && n_layers <= MAX_LAYERS, PTK_ERR_SHAPE, "gcn_stack_fwd: bad sizes");
    const int64_t M = B * Nv;
    PTK_REQUIRE(algo >= 0 && algo <= 2, PTK_ERR_SHAPE, "gcn_stack_fwd: algo must be 0, 1 or 2");
    const FwdPlan p = plan_fwd(M, n_layers, widths, Ls, algo);
    PTK_REQUIRE(workspace && workspace_bytes >= p.total, PTK_ERR_WORKSPACE, "gcn_stack_fwd: workspace too small");
    PTK_REQUIRE(aligned16(workspace), PTK_ERR_ALIGN, "gcn_stack_fwd: workspace must be 16-byte aligned");
    char *ws = reinterpret_cast<char *>(workspace);
    float *head = reinterpret_cast<float *>(ws + p.head_off);
    float *Hbuf = reinterpret_cast<float *>(ws + p.h_off);
    void *lin = ws + p.lin_off;
    const size_t lin_bytes = p.lin_bytes;
    cudaStream_t st = as_stream(stream);
    // tensor-core layers: all weight splits of the pass in one launch (instead of one per layer)
    static thread_local float *b_hi[MAX_LAYERS], *b_lo[MAX_LAYERS];
    static thread_local bool tc[MAX_LAYERS];
    int *flags = reinterpret_cast<int *>(ws + p.flags_off);
    float *part = reinterpret_cast<float *>(ws + p.part_off);
    if (algo != PTK_GEMM_FFMA) {
        static thread_local const float *src[MAX_LAYERS];
        static thread_local float *hi[MAX_LAYERS], *lo[MAX_LAYERS];
        static thread_local int rows[MAX_LAYERS], cols[MAX_LAYERS];
        int cnt = 0;
        size_t off = p.split_off;
        for (int l = 0; l < n_layers; ++l) {
            const int64_t K = widths[l], N = widths[l + 1];
            b_hi[l] = reinterpret_cast<float *>(ws + off);
            b_lo[l] = b_hi[l] + (size_t)K * N;
            off += align_up(2 * sizeof(float) * (size_t)K * (size_t)N);
            // layer inputs and Hbuf are 16-byte aligned whenever X and acts are (checked per layer below)
            tc[l] = W[l] && tf32x3_eligible(l == 0 ? (const void *)X : (const void *)acts[l - 1], Hbuf, M, K, N);
            if (tc[l]) {
                src[cnt] = W[l]; hi[cnt] = b_hi[l]; lo[cnt] = b_lo[l]; rows[cnt] = (int)K; cols[cnt] = (int)N;
                ++cnt;
            }
        }
        if (cnt > 0) {
            int rc = tf32x3_presplit_batched(cnt, src, rows, cols, /*transpose=*/1, hi, lo, flags,
                                             n_layers * tf32x3_max_ctas(), st);
            if (rc) return rc;
        }
    } else {
        for (int l = 0; l < n_layers; ++l) tc[l] = false;
    }
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
            if (tc[l])
                rc = gemm_tf32x3_presplit(in, b_hi[l], b_lo[l], nullptr, nullptr, M, K, N, Hbuf, nullptr, part,
                                          flags + (size_t)l * tf32x3_max_ctas(), st);
            else
                rc = ptk_gcn_linear_fwd(in, W[l], M, K, N, Hbuf, algo == PTK_GEMM_TF32X3 ? algo : PTK_GEMM_FFMA, lin,
                                        lin_bytes, stream);
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
    PTK_NVTX("ptk_gcn_stack_bwd");
    PTK_REQUIRE(graph_t && widths && Ls && relus && X && W && acts && gout && gW && gb && need_gb, PTK_ERR_SHAPE,
                "gcn_stack_bwd: null pointer");
    PTK_REQUIRE(B > 0 && Nv > 0 && n_layers > 0 && n_layers <= MAX_LAYERS, PTK_ERR_SHAPE, "gcn_stack_bwd: bad sizes");
    const int n = n_layers;
    const int64_t M = B * Nv;
    BwdPlan p = plan_bwd(M, n, widths, Ls, need_gb, batch_bias);
    PTK_REQUIRE(workspace && workspace_bytes >= p.total, PTK_ERR_WORKSPACE, "gcn_stack_bwd: workspace too small");
    PTK_REQUIRE(aligned16(workspace), PTK_ERR_ALIGN, "gcn_stack_bwd: workspace must be 16-byte aligned");
    PTK_REQUIRE(algo_dgrad >= 0 && algo_dgrad <= 2 && algo_wgrad >= 0 && algo_wgrad <= 2, PTK_ERR_SHAPE,
                "gcn_stack_bwd: algo must be 0, 1 or 2");
    char *ws = reinterpret_cast<char *>(workspace);
    float *gH = reinterpret_cast<float *>(ws + p.gh_off);
    float *tmp = reinterpret_cast<float *>(ws + p.tmp_off);
    void *lin = ws + p.lin_off;
    void *bg = ws + p.bg_off;
    float *slab = reinterpret_cast<float *>(ws + p.slab_off);
    cudaStream_t st = as_stream(stream);
    int *flags = reinterpret_cast<int *>(ws + p.flags_off);
    float *part = reinterpret_cast<float *>(ws + p.part_off);
    uint32_t *mask_ws = reinterpret_cast<uint32_t *>(ws + p.mask_off);
    // tensor-core dgrad: gX = gH . W^T reduces over the layer's OUTPUT width; W (K_in x N_out) row-major is the K-major
    // operand as stored.  All splits of the pass in one launch.
    static thread_local float *b_hi[MAX_LAYERS], *b_lo[MAX_LAYERS];
    static thread_local bool tc_d[MAX_LAYERS];
    {
        static thread_local const float *src[MAX_LAYERS];
        static thread_local float *hi[MAX_LAYERS], *lo[MAX_LAYERS];
        static thread_local int rows[MAX_LAYERS], cols[MAX_LAYERS];
        int cnt = 0;
        size_t off = p.split_off;
        for (int l = 0; l < n; ++l) {
            const int64_t K = widths[l], N = widths[l + 1];
            b_hi[l] = reinterpret_cast<float *>(ws + off);
            b_lo[l] = b_hi[l] + (size_t)K * N;
            off += align_up(2 * sizeof(float) * (size_t)K * (size_t)N);
            const bool wanted = l > 0 || gX;
            const void *act_in = l == 0 ? (const void *)X : (const void *)acts[l - 1];
            tc_d[l] = wanted && algo_dgrad != PTK_GEMM_FFMA && W[l] && tf32x3_eligible(gH, l == 0 ? (void *)gX : (void *)tmp, M, N, K) &&
                      aligned16(act_in);
            if (tc_d[l]) {
                src[cnt] = W[l]; hi[cnt] = b_hi[l]; lo[cnt] = b_lo[l]; rows[cnt] = (int)K; cols[cnt] = (int)N;
                ++cnt;
            }
        }
        if (cnt > 0) {
            int rc0 = tf32x3_presplit_batched(cnt, src, rows, cols, /*transpose=*/0, hi, lo, flags, n * tf32x3_max_ctas(), st);
            if (rc0) return rc0;
        }
    }
    ReduceBatch rb;
    int rb_n = 0;
    auto flush_reduce = [&]() -> int {
        if (rb_n == 0) return PTK_OK;
        long long mx = 0;
        for (int i = 0; i < rb_n; ++i)
            if (rb.elems[i] > mx) mx = rb.elems[i];
        for (int i = rb_n; i < RB_MAX; ++i) { rb.part[i] = rb.part[0]; rb.out[i] = rb.out[0]; rb.ns[i] = 0; rb.elems[i] = 0; }
        launch_pdl(splitk_reduce_batched_kernel, dim3((unsigned)ceil_div(mx, 256), (unsigned)rb_n), dim3(256), 0, st, rb);
        PTK_CHECK_LAUNCH();
        rb_n = 0;
        return PTK_OK;
    };
    size_t wg_off = p.wg_off;
    // the batched bias gradients are written as one (group_n, width) matrix: the group's gb pointers must be that
    int first = -1;
    static thread_local int slot_of[MAX_LAYERS];
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
        const size_t wg_bytes = align_up(wgrad_tf32x3_workspace_bytes(M, K, N));
        if (gW[l]) {
            if (algo_wgrad != PTK_GEMM_FFMA && wgrad_tf32x3_eligible(act_in, gH, M, K, N)) {
                // split-M partial sums stay in this layer's own region; one batched reduction ends the pass
                float *wpart = nullptr;
                int ns = 0;
                rc = wgrad_tf32x3(act_in, gH, M, K, N, ws + wg_off, wg_bytes, &wpart, &ns, st);
                if (rc) return rc;
                rb.part[rb_n] = wpart; rb.out[rb_n] = gW[l]; rb.ns[rb_n] = ns; rb.elems[rb_n] = (long long)K * N;
                if (++rb_n == RB_MAX) {
                    rc = flush_reduce();
                    if (rc) return rc;
                }
            } else {
                rc = ptk_gcn_linear_wgrad(act_in, gH, M, K, N, gW[l], algo_wgrad == PTK_GEMM_TF32X3 ? algo_wgrad : PTK_GEMM_FFMA,
                                          lin, p.lin_bytes, stream);
                if (rc) return rc;
            }
        }
        wg_off += wg_bytes;
        if (l > 0 || gX) {
            const bool masked = l > 0 && relus[l - 1];
            float *dst = l == 0 ? gX : (slot(l - 1) >= 0 ? slab + (size_t)slot(l - 1) * M * p.group_w : tmp);
            const uint32_t *bits = (masked && x_bits) ? x_bits[l] : nullptr;
            if (tc_d[l] && aligned16(dst))
                rc = gemm_tf32x3_presplit(gH, b_hi[l], b_lo[l], masked ? act_in : nullptr, bits, M, /*reduce over*/ N,
                                          /*outputs*/ K, dst, mask_ws, part, flags + (size_t)l * tf32x3_max_ctas(), st);
            else
                rc = ptk_gcn_linear_dgrad(gH, W[l], masked ? act_in : nullptr, bits, M, K, N, dst,
                                          algo_dgrad == PTK_GEMM_TF32X3 ? algo_dgrad : PTK_GEMM_FFMA, lin, p.lin_bytes,
                                          stream);
            if (rc) return rc;
            g = dst;
        }
    }
    rc = flush_reduce();
    if (rc) return rc;
    if (p.group_n > 0) {
        rc = ptk_gcn_bias_grad_batched(slab, p.group_n, M, p.group_w, p.group_L, gb[first], bg, p.bg_bytes, stream);
        if (rc) return rc;
    }
    return PTK_OK;
}
