// Helpers shared by the tile kernels of the GCN aggregation (gcn_aggregate.cu, gcn_aggregate_union.cu).
#pragma once
#include "ptk_common.cuh"

namespace ptk {

constexpr int AG_THREADS = 256;
constexpr int AG_WARPS = AG_THREADS / 32;
constexpr int HUB_DEG = 128;
constexpr int AT_STRIP = 128;  // staged neighbours per warp (= HUB_DEG: a non-hub row fits)

template <int NG>
__device__ __forceinline__ void strip_gather(const uint32_t *__restrict__ s_off, const float *__restrict__ s_w,
                                             int n4, const char *const (&base)[NG], float (&acc)[NG][4]) {
    // n4: staged entries, a multiple of 4 (padding has weight 0 and a valid offset).  base[n] already
    // includes the lane's channel-group offset; lanes beyond the last aggregated group are clamped onto it
    // (same cache lines, result discarded), which keeps the loop free of divergent branches.
    int k = 0;
    for (; k + 8 <= n4; k += 8) {
        const uint4 o0 = *reinterpret_cast<const uint4 *>(s_off + k), o1 = *reinterpret_cast<const uint4 *>(s_off + k + 4);
        const float4 w0 = *reinterpret_cast<const float4 *>(s_w + k), w1 = *reinterpret_cast<const float4 *>(s_w + k + 4);
        const uint32_t off[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
        const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int n = 0; n < NG; ++n) {
            float4 a[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) a[u] = *reinterpret_cast<const float4 *>(base[n] + off[u]);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                acc[n][0] = fmaf(w[u], a[u].x, acc[n][0]); acc[n][1] = fmaf(w[u], a[u].y, acc[n][1]);
                acc[n][2] = fmaf(w[u], a[u].z, acc[n][2]); acc[n][3] = fmaf(w[u], a[u].w, acc[n][3]);
            }
        }
    }
    if (k < n4) {
        const uint4 o0 = *reinterpret_cast<const uint4 *>(s_off + k);
        const float4 w0 = *reinterpret_cast<const float4 *>(s_w + k);
        const uint32_t off[4] = {o0.x, o0.y, o0.z, o0.w};
        const float w[4] = {w0.x, w0.y, w0.z, w0.w};
#pragma unroll
        for (int n = 0; n < NG; ++n) {
            float4 a[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) a[u] = *reinterpret_cast<const float4 *>(base[n] + off[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                acc[n][0] = fmaf(w[u], a[u].x, acc[n][0]); acc[n][1] = fmaf(w[u], a[u].y, acc[n][1]);
                acc[n][2] = fmaf(w[u], a[u].z, acc[n][2]); acc[n][3] = fmaf(w[u], a[u].w, acc[n][3]);
            }
        }
    }
}

// Stage entries [e0, e0+cnt) of a row into the warp's strip (cnt <= AT_STRIP), padded to a multiple of 4.
__device__ __forceinline__ int strip_stage(const int32_t *__restrict__ col, const float *__restrict__ val, int e0,
                                           int cnt, uint32_t row_bytes, uint32_t *s_off, float *s_w) {
    const int lane = threadIdx.x & 31;
    const int n4 = (cnt + 3) & ~3;
    __syncwarp();
    for (int e = lane; e < n4; e += 32) {
        const bool real = e < cnt;
        s_off[e] = (uint32_t)col[e0 + (real ? e : 0)] * row_bytes;
        s_w[e] = real ? val[e0 + e] : 0.f;
    }
    __syncwarp();
    return n4;
}

// Epilogue of one output row: aggregated groups (+bias, boundary group mixes in the pass-through channels).
template <int NG>
__device__ __forceinline__ void row_epilogue(const float (&acc)[NG][4], const bool (&on)[NG], const float *__restrict__ self,
                                             float *__restrict__ o, const float *__restrict__ bias, int L, int relu) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int n = 0; n < NG; ++n) {
        if (!on[n]) continue;
        const int c0 = (lane + 32 * n) * 4;
        float r[4] = {acc[n][0], acc[n][1], acc[n][2], acc[n][3]};
        if (bias) {
            const float4 bv = *reinterpret_cast<const float4 *>(bias + c0);
            r[0] += bv.x; r[1] += bv.y; r[2] += bv.z; r[3] += bv.w;
        }
        if (c0 + 4 > L) {  // boundary group: channels >= L pass through
            const float4 s = *reinterpret_cast<const float4 *>(self + c0);
            const float sv[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (c0 + k >= L) r[k] = sv[k];
        }
        if (relu) {
#pragma unroll
            for (int k = 0; k < 4; ++k) r[k] = fmaxf(r[k], 0.f);
        }
        __stcs(reinterpret_cast<float4 *>(o + c0), make_float4(r[0], r[1], r[2], r[3]));
    }
}

__device__ __forceinline__ float4 relu4(float4 s, int relu) {
    if (relu) {
        s.x = fmaxf(s.x, 0.f); s.y = fmaxf(s.y, 0.f); s.z = fmaxf(s.z, 0.f); s.w = fmaxf(s.w, 0.f);
    }
    return s;
}

// Hub rows with a shared neighbour set ("common set").  The touch-chart centre vertices are each linked to
// ALL boundary vertices (utils.py:126-128): the 5 (finger) or 20 (grasp) hub rows of the fused graph read
// the same ~1150 neighbour rows.  When the host finds such a set S with val[h,j] = alpha[h] * cw[j] on it
// (graph.py), one hub CTA per batch element computes  y* = sum_{j in S} cw[j] x_j  ONCE, and every hub row
// becomes  alpha[h] * y* + (its few remaining neighbours, kept in the reduced CSR)  -- 1/n_hubs of the
// gather traffic.  Rows flagged in row_skip are left to the hub CTAs by the tile CTAs.
struct AggHubs {
    const int32_t *hubs;        // hub row ids (n_hubs)
    int n_hubs;
    const int32_t *common_col;  // common set (n_common), may be NULL: every hub row is gathered on its own
    const float *common_w;
    int n_common;
    const float *alpha;         // per hub row scale of y* (n_hubs)
    const uint8_t *row_skip;    // Nv flags, may be NULL: rows of degree > HUB_DEG are the hub rows
};

}  // namespace ptk
