// Vertex-feature front of the deformation network, fused: positions -> GCN layer-0 input in ONE launch.
//
// Replaces, per deformation iteration (pterotactyl/reconstruction/vision/model.py:229-236, 261-267, 274-279):
//   Positional_Encoder.forward  (model.py:393-399): NeRF embedding (20 sin/cos launches, 20 multiplies, 2 cat),
//                                                   Linear(63, s/4) + ReLU, Linear(s/4, s/2) + ReLU, Linear(s/2, s)
//   Mask_Encoder.forward        (model.py:410-414): Embedding(4, s) lookup of the per-vertex mask token
//   vertex_features = positional_features + mask_features [+ img_features]   (model.py:234-236)
// i.e. ~50 launches, three library GEMMs and four (M x s) round trips through HBM.  Here a CTA owns 64 vertices:
// it evaluates their embedding into shared memory, runs the three layers with the activations resident in shared
// memory (k-major, so a thread reads 4 vertices of one k with a single LDS.128) while the weights stream from L2 in
// 16 x 128 tiles, and the last epilogue adds bias, the mask token's embedding row and (optionally) the pooled image
// features before the only global store.  FP32 FMA chains in ascending k (the arithmetic of a scalar loop), accurate
// sinf / cosf as in nerf_embed.cu.  In training the two hidden activations are also written out (the backward --
// ops._VertexFront -- needs them for the weight gradients and ReLU masks).
#include "ptk_common.cuh"

namespace ptk {

constexpr int VF_TM = 64;        // vertices per CTA
constexpr int VF_THREADS = 256;  // 16 row groups (4 vertices) x 16 column groups (2 x 4 outputs)
constexpr int VF_NT = 128;       // output columns per weight tile
constexpr int VF_KT = 16;        // reduction steps per weight tile
constexpr int VF_WP = VF_NT + 4;  // pitch of a staged weight row (floats): the transposing stores conflict 2-way, not 4-way
constexpr int VF_EMB = 63, VF_EMB_P = 64;
constexpr int VF_MAX_H1 = 112, VF_MAX_H2 = 224;  // hidden widths of input_size <= 448 (+ padding to 16)

__device__ __forceinline__ float vf_scale(int i) {
    // fl32 of the Python doubles np.pi (i = 0) and np.pi * 2 * i  (model.py:385-389; nerf_embed.cu)
    return i == 0 ? (float)3.141592653589793 : (float)((3.141592653589793 * 2.0) * (double)i);
}

struct VFParams {
    const float *pos, *mask;
    const float *w1, *b1, *w2, *b2, *w3, *b3;  // nn.Linear layout: (out, in) row-major
    const float *emb, *add;
    long long M;
    int h1, h2, width;
    float *out, *h1_save, *h2_save;
};

// One layer over the CTA's 64 vertices:  res[m][n] = sum_k in_t[k][m] * W[n][k]  (+ epilogue), n in [0, Nout).
//   in_t : shared, k-major [>= ceil16(K)][VF_TM]; rows beyond K hold finite values (their weights are staged as 0)
//   LAST = false: out_t[n][m] = relu(res + bias[n])  (shared, k-major for the next layer), optional global copy
//   LAST = true : out[m][n]   = res + bias[n] + emb[token[m]][n] + add[m][n]
template <bool LAST>
__device__ __forceinline__ void vf_layer(const float *__restrict__ W, const float *__restrict__ bias, int K, int Nout,
                                         const float *in_t, float *out_t, float (*ws)[VF_KT][VF_WP], long long m0,
                                         long long M, float *__restrict__ gout, const float *__restrict__ emb,
                                         const int *s_tok, const float *__restrict__ add) {
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    // Weight staging.  K % 4 == 0 (every layer but the first): 4 lanes bring the 16 k of one output row as float4 (64
    // contiguous bytes per row, 2 rows per thread) and store them transposed; otherwise (K = 63) scalar loads,
    // thread = (row, 8 consecutive k).
    const bool vec_w = (K & 3) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0;
    const int vrow = tid >> 2, vk = (tid & 3) * 4;
    const int sn = tid & (VF_NT - 1), sk = (tid >> 7) * 8;
    const int nkt = (K + VF_KT - 1) / VF_KT;
    const bool vec_out = (Nout & 3) == 0;
    for (int n0 = 0; n0 < Nout; n0 += VF_NT) {
        // accumulators as packed FP32x2 pairs along n: fma.rn.f32x2 = the same IEEE FMA per element, half the issue slots
        unsigned long long acc2[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc2[i][j] = 0ull;
        float stg[8];
        auto fetch = [&](int kt) {
            if (vec_w) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int n = n0 + vrow + 64 * h, k = kt * VF_KT + vk;
                    const float4 v = (n < Nout && k < K) ? __ldg(reinterpret_cast<const float4 *>(W + (size_t)n * K + k))
                                                         : make_float4(0.f, 0.f, 0.f, 0.f);
                    stg[4 * h] = v.x; stg[4 * h + 1] = v.y; stg[4 * h + 2] = v.z; stg[4 * h + 3] = v.w;
                }
            } else {
                const int n = n0 + sn;
                const float *src = W + (size_t)n * K + kt * VF_KT + sk;
#pragma unroll
                for (int j = 0; j < 8; ++j) stg[j] = (n < Nout && kt * VF_KT + sk + j < K) ? __ldg(src + j) : 0.f;
            }
        };
        auto stash = [&](int buf) {
            if (vec_w) {
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int j = 0; j < 4; ++j) ws[buf][vk + j][vrow + 64 * h] = stg[4 * h + j];
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) ws[buf][sk + j][sn] = stg[j];
            }
        };
        fetch(0);
        __syncthreads();  // the previous user of ws / the producer of in_t is done
        stash(0);
        __syncthreads();
        for (int kt = 0; kt < nkt; ++kt) {
            const int buf = kt & 1;
            if (kt + 1 < nkt) fetch(kt + 1);
#pragma unroll
            for (int kk = 0; kk < VF_KT; ++kk) {
                const float4 a = *reinterpret_cast<const float4 *>(in_t + (size_t)(kt * VF_KT + kk) * VF_TM + ty * 4);
                const ulonglong2 b0 = *reinterpret_cast<const ulonglong2 *>(&ws[buf][kk][tx * 4]);
                const ulonglong2 b1 = *reinterpret_cast<const ulonglong2 *>(&ws[buf][kk][64 + tx * 4]);
                const float av[4] = {a.x, a.y, a.z, a.w};
                const unsigned long long bv[4] = {b0.x, b0.y, b1.x, b1.y};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    unsigned long long a2;
                    asm("mov.b64 %0, {%1, %1};" : "=l"(a2) : "f"(av[i]));
#pragma unroll
                    for (int j = 0; j < 4; ++j) asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2[i][j]) : "l"(a2), "l"(bv[j]));
                }
            }
            if (kt + 1 < nkt) {
                stash(buf ^ 1);
                __syncthreads();
            }
        }
        float acc[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc[i][2 * j] = __uint_as_float((unsigned)acc2[i][j]);
                acc[i][2 * j + 1] = __uint_as_float((unsigned)(acc2[i][j] >> 32));
            }
        // epilogue: this thread holds vertices ty*4 .. +3, columns n0 + tx*4 .. +3 and n0 + 64 + tx*4 .. +3
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int nb = n0 + h * 64 + tx * 4;
            if (nb >= Nout) continue;
            float bb[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) bb[j] = nb + j < Nout ? __ldg(bias + nb + j) : 0.f;
            if (!LAST) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (nb + j >= Nout) continue;
                    float4 v;
                    v.x = fmaxf(acc[0][h * 4 + j] + bb[j], 0.f);
                    v.y = fmaxf(acc[1][h * 4 + j] + bb[j], 0.f);
                    v.z = fmaxf(acc[2][h * 4 + j] + bb[j], 0.f);
                    v.w = fmaxf(acc[3][h * 4 + j] + bb[j], 0.f);
                    *reinterpret_cast<float4 *>(out_t + (size_t)(nb + j) * VF_TM + ty * 4) = v;
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const long long m = m0 + ty * 4 + i;
                if (m >= M) continue;
                float r[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) r[j] = acc[i][h * 4 + j] + bb[j];
                if (LAST) {
                    if (emb) {
                        const float *e = emb + (size_t)s_tok[ty * 4 + i] * Nout + nb;
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (nb + j < Nout) r[j] += __ldg(e + j);
                    }
                    if (add) {
                        const float *e = add + (size_t)m * Nout + nb;
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (nb + j < Nout) r[j] += __ldg(e + j);
                    }
                } else {
                    if (!gout) continue;
#pragma unroll
                    for (int j = 0; j < 4; ++j) r[j] = fmaxf(r[j], 0.f);
                }
                float *dst = gout + (size_t)m * Nout + nb;
                if (vec_out) {
                    *reinterpret_cast<float4 *>(dst) = make_float4(r[0], r[1], r[2], r[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (nb + j < Nout) dst[j] = r[j];
                }
            }
        }
    }
}

__global__ void __launch_bounds__(VF_THREADS, 2)
vertex_front_fwd_kernel(const VFParams p) {
    extern __shared__ __align__(16) float vf_smem[];
    // [act A: embedding, later hidden 2][act B: hidden 1][weight tiles 2 x 16 x 128][tokens]
    float *act_a = vf_smem;                                   // max(64, ceil16(h2)) x 64
    float *act_b = act_a + (size_t)VF_MAX_H2 * VF_TM;         // ceil16(h1) x 64
    float(*ws)[VF_KT][VF_WP] = reinterpret_cast<float(*)[VF_KT][VF_WP]>(act_b + (size_t)VF_MAX_H1 * VF_TM);
    int *s_tok = reinterpret_cast<int *>(&ws[2][0][0]);
    const int tid = threadIdx.x;
    const long long m0 = (long long)blockIdx.x * VF_TM;
    pdl_launch_dependents();
    pdl_wait();
    // finite values everywhere a padded k row can be read
    for (int e = tid; e < (VF_MAX_H2 + VF_MAX_H1) * VF_TM; e += VF_THREADS) vf_smem[e] = 0.f;
    if (tid < VF_TM) {
        int t = 0;
        if (p.mask && m0 + tid < p.M) {
            t = (int)p.mask[m0 + tid];  // .long(): truncation (model.py:413)
            t = t < 0 ? 0 : (t > 3 ? 3 : t);
        }
        s_tok[tid] = t;
    }
    __syncthreads();
    // NeRF embedding, k-major: act_a[c][m], c = 6 i + {0,1,2: sin | 3,4,5: cos} of (s_i * p), c = 60..62: p
    for (int e = tid; e < VF_EMB * VF_TM; e += VF_THREADS) {
        const int c = e / VF_TM, m = e - c * VF_TM;
        float v = 0.f;
        if (m0 + m < p.M) {
            if (c >= 60) {
                v = p.pos[(m0 + m) * 3 + (c - 60)];
            } else {
                const int i = c / 6, r = c - 6 * i;
                const float arg = __fmul_rn(vf_scale(i), p.pos[(m0 + m) * 3 + (r < 3 ? r : r - 3)]);
                v = r < 3 ? sinf(arg) : cosf(arg);
            }
        }
        act_a[e] = v;
    }
    // (vf_layer starts with a barrier)
    vf_layer<false>(p.w1, p.b1, VF_EMB, p.h1, act_a, act_b, ws, m0, p.M, p.h1_save, nullptr, nullptr, nullptr);
    vf_layer<false>(p.w2, p.b2, p.h1, p.h2, act_b, act_a, ws, m0, p.M, p.h2_save, nullptr, nullptr, nullptr);
    vf_layer<true>(p.w3, p.b3, p.h2, p.width, act_a, nullptr, ws, m0, p.M, p.out, p.emb, s_tok, p.add);
}

// ---- gradients of the last bias and of the embedding table: column sums of g per mask token.
// stage 1: part[slab][t][n] = sum over the slab's rows with token t of g[m][n]; stage 2: fixed-order sum over slabs.
constexpr int VC_ROWS = 256;

// block (128 columns, 4 row groups): row group y sums rows r0 + 64 y .. + 63 of the slab (eight rows' loads in flight per
// thread: the kernel is latency-bound), the groups are combined through shared memory in group order -- a fixed
// summation order, independent of scheduling.
constexpr int VC_GROUPS = 4;
__global__ void __launch_bounds__(128 * VC_GROUPS)
vertex_front_colsum1_kernel(const float *__restrict__ g, const float *__restrict__ mask, long long M, int N,
                            float *__restrict__ part) {
    pdl_wait();
    __shared__ float sacc[VC_GROUPS][4][128];
    const int n = blockIdx.x * 128 + threadIdx.x, grp = threadIdx.y;
    const long long s0 = (long long)blockIdx.y * VC_ROWS;
    const long long r0 = s0 + (long long)grp * (VC_ROWS / VC_GROUPS);
    long long r1 = r0 + VC_ROWS / VC_GROUPS;
    r1 = r1 < M ? r1 : M;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (n < N) {
        for (long long m0 = r0; m0 < r1; m0 += 8) {
            float v[8];
            int t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const long long m = m0 + u < r1 ? m0 + u : r1 - 1;
                t[u] = mask ? (int)__ldg(mask + m) : 0;
                v[u] = __ldg(g + (size_t)m * N + n);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (m0 + u >= r1) break;
                const int tt = t[u] < 0 ? 0 : (t[u] > 3 ? 3 : t[u]);
                acc[0] += tt == 0 ? v[u] : 0.f;
                acc[1] += tt == 1 ? v[u] : 0.f;
                acc[2] += tt == 2 ? v[u] : 0.f;
                acc[3] += tt == 3 ? v[u] : 0.f;
            }
        }
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) sacc[grp][t][threadIdx.x] = acc[t];
    __syncthreads();
    if (grp == 0 && n < N) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            float a = sacc[0][t][threadIdx.x];
#pragma unroll
            for (int y = 1; y < VC_GROUPS; ++y) a += sacc[y][t][threadIdx.x];
            part[((size_t)blockIdx.y * 4 + t) * N + n] = a;
        }
    }
}

__global__ void __launch_bounds__(256)
vertex_front_colsum2_kernel(const float *__restrict__ part, int slabs, int N, float *__restrict__ sums) {
    pdl_wait();
    const int e = blockIdx.x * 256 + threadIdx.x;  // (t, n)
    if (e >= 4 * N) return;
    float acc = 0.f;
    for (int s = 0; s < slabs; ++s) acc += part[(size_t)s * 4 * N + e];
    sums[e] = acc;
}

static unsigned long long g_vf_optin = 0ull;

}  // namespace ptk

using namespace ptk;

extern "C" int ptk_vertex_front_fwd(const float *positions, const float *mask, const float *w1, const float *b1,
                                    const float *w2, const float *b2, const float *w3, const float *b3,
                                    const float *emb, const float *add, int64_t M, int32_t h1, int32_t h2,
                                    int32_t width, float *out, float *h1_save, float *h2_save, ptk_stream_t stream) {
    PTK_NVTX("ptk_vertex_front_fwd");
    PTK_REQUIRE(M >= 0 && M < (1LL << 31), PTK_ERR_SHAPE, "vertex_front_fwd: bad M = %lld", (long long)M);
    PTK_REQUIRE(h1 >= 1 && h1 <= VF_MAX_H1 && h2 >= 1 && h2 <= VF_MAX_H2 && width >= 1, PTK_ERR_SHAPE,
                "vertex_front_fwd: hidden widths (%d, %d) outside (1..%d, 1..%d)", h1, h2, VF_MAX_H1, VF_MAX_H2);
    if (M == 0) return PTK_OK;
    PTK_REQUIRE(positions && w1 && b1 && w2 && b2 && w3 && b3 && out, PTK_ERR_SHAPE, "vertex_front_fwd: null pointer");
    PTK_REQUIRE((width % 4) != 0 || (((uintptr_t)out) % 16) == 0, PTK_ERR_ALIGN, "vertex_front_fwd: out must be 16-byte aligned");
    PTK_REQUIRE(!h1_save || (h1 % 4) != 0 || (((uintptr_t)h1_save) % 16) == 0, PTK_ERR_ALIGN, "vertex_front_fwd: h1_save alignment");
    PTK_REQUIRE(!h2_save || (h2 % 4) != 0 || (((uintptr_t)h2_save) % 16) == 0, PTK_ERR_ALIGN, "vertex_front_fwd: h2_save alignment");
    VFParams p;
    p.pos = positions; p.mask = mask; p.w1 = w1; p.b1 = b1; p.w2 = w2; p.b2 = b2; p.w3 = w3; p.b3 = b3;
    p.emb = emb; p.add = add; p.M = M; p.h1 = h1; p.h2 = h2; p.width = width; p.out = out;
    p.h1_save = h1_save; p.h2_save = h2_save;
    const size_t smem = sizeof(float) * ((size_t)(VF_MAX_H2 + VF_MAX_H1) * VF_TM + 2 * VF_KT * VF_WP) + sizeof(int) * VF_TM;
    int dev = 0;
    PTK_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev >= 64 || !((g_vf_optin >> dev) & 1ull)) {
        PTK_CHECK_CUDA(cudaFuncSetAttribute(vertex_front_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev < 64) g_vf_optin |= 1ull << dev;
    }
    launch_pdl(vertex_front_fwd_kernel, dim3((unsigned)ceil_div(M, VF_TM)), dim3(VF_THREADS), smem, as_stream(stream), p);
    PTK_CHECK_LAUNCH();
    return PTK_OK;
}

extern "C" size_t ptk_vertex_front_colsum_workspace_bytes(int64_t M, int32_t width) {
    if (M <= 0 || width <= 0) return 0;
    return sizeof(float) * (size_t)ceil_div(M, VC_ROWS) * 4 * (size_t)width;
}

extern "C" int ptk_vertex_front_colsum(const float *g, const float *mask, int64_t M, int32_t width, float *sums,
                                       void *workspace, size_t workspace_bytes, ptk_stream_t stream) {
    PTK_REQUIRE(M > 0 && width > 0 && g && sums, PTK_ERR_SHAPE, "vertex_front_colsum: bad argument");
    PTK_REQUIRE(workspace && workspace_bytes >= ptk_vertex_front_colsum_workspace_bytes(M, width), PTK_ERR_WORKSPACE,
                "vertex_front_colsum: workspace too small");
    const int slabs = (int)ceil_div(M, VC_ROWS);
    PTK_REQUIRE(slabs <= 65535, PTK_ERR_SHAPE, "vertex_front_colsum: M too large");
    float *part = reinterpret_cast<float *>(workspace);
    launch_pdl(vertex_front_colsum1_kernel, dim3((unsigned)ceil_div(width, 128), (unsigned)slabs), dim3(128, VC_GROUPS), 0,
               as_stream(stream), g, mask, (long long)M, (int)width, part);
    PTK_CHECK_LAUNCH();
    launch_pdl(vertex_front_colsum2_kernel, dim3((unsigned)ceil_div(4 * width, 256)), dim3(256), 0, as_stream(stream),
               (const float *)part, slabs, (int)width, sums);
    PTK_CHECK_LAUNCH();
    return PTK_OK;
}
