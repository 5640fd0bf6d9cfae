// GCN vertex aggregation (CSR gather) for sm_100a.
//
// Replaces the dense `torch.matmul(adj, features[:, :, :length])` + cat + bias + activation of
// GCN_layer.forward (pterotactyl/reconstruction/vision/model.py:354-363): the reference multiplies a
// 99 %-zero (Nv x Nv) matrix into (B,Nv,L); here each output row gathers its deg(i) neighbour rows.
//
//   out[b,i,c] = act( sum_{e in row i} val[e] * in[b,col[e],c] + bias[c] )   c <  L
//   out[b,i,c] = act( in[b,i,c] )                                             c >= L
//
// Kernels in this file: gcn_aggregate_tile_kernel (the fast path, see the comment above it: row structure
// staged once per warp and replayed over a batch group, hub rows factored over their common neighbour set),
// gcn_aggregate_narrow_kernel (<= 8 channels, the 3-channel output layer) and gcn_aggregate_kernel, the
// generic first version described next, kept for shapes the other two do not take (C % 4 != 0, L = 0):
//
// HBM/L2-bound: one warp per output row, 128-bit loads along the channel dimension (rows are
// C*4 bytes = 1200 B for C=300: 16-byte aligned), the column indices of a row are fetched 32 at a
// time by the warp and broadcast with shuffles, neighbour loads are issued 4 deep.  Hub rows
// (degree > HUB_DEG: the touch-chart centre vertices, degree 1153 -- utils.py:95-98,126-128) are
// processed by a whole CTA (8 warps split the neighbour list, shared-memory reduction) so that
// they do not become the tail of the launch.
#include <stdlib.h>

#include "gcn_aggregate_common.cuh"

namespace ptk {

// Gather for one row restricted to the warp `part` of `nparts` (nparts = 1: the whole row).
template <bool VEC>
__device__ __forceinline__ void row_gather(const int32_t *__restrict__ col, const float *__restrict__ val,
                                           int beg, int end, const float *__restrict__ inb, int C,
                                           int v, bool active, float acc[4]) {
    const int lane = threadIdx.x & 31;
    for (int e0 = beg; e0 < end; e0 += 32) {
        const int cnt = min(32, end - e0);
        int my_col = 0;
        float my_val = 0.f;
        if (lane < cnt) {
            my_col = col[e0 + lane];
            my_val = val[e0 + lane];
        }
        int k = 0;
        for (; k + 8 <= cnt; k += 8) {  // 8 neighbour rows in flight per lane
            int j[8];
            float w[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                j[u] = __shfl_sync(0xffffffffu, my_col, k + u);
                w[u] = __shfl_sync(0xffffffffu, my_val, k + u);
            }
            if (active) {
                if (VEC) {
                    float4 a[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) a[u] = *reinterpret_cast<const float4 *>(inb + (size_t)j[u] * C + v * 4);
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        acc[0] = fmaf(w[u], a[u].x, acc[0]); acc[1] = fmaf(w[u], a[u].y, acc[1]);
                        acc[2] = fmaf(w[u], a[u].z, acc[2]); acc[3] = fmaf(w[u], a[u].w, acc[3]);
                    }
                } else {
                    float a[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) a[u] = inb[(size_t)j[u] * C + v];
#pragma unroll
                    for (int u = 0; u < 8; ++u) acc[0] = fmaf(w[u], a[u], acc[0]);
                }
            }
        }
        for (; k + 4 <= cnt; k += 4) {
            int j0 = __shfl_sync(0xffffffffu, my_col, k + 0), j1 = __shfl_sync(0xffffffffu, my_col, k + 1);
            int j2 = __shfl_sync(0xffffffffu, my_col, k + 2), j3 = __shfl_sync(0xffffffffu, my_col, k + 3);
            float w0 = __shfl_sync(0xffffffffu, my_val, k + 0), w1 = __shfl_sync(0xffffffffu, my_val, k + 1);
            float w2 = __shfl_sync(0xffffffffu, my_val, k + 2), w3 = __shfl_sync(0xffffffffu, my_val, k + 3);
            if (active) {
                if (VEC) {
                    float4 a0 = *reinterpret_cast<const float4 *>(inb + (size_t)j0 * C + v * 4);
                    float4 a1 = *reinterpret_cast<const float4 *>(inb + (size_t)j1 * C + v * 4);
                    float4 a2 = *reinterpret_cast<const float4 *>(inb + (size_t)j2 * C + v * 4);
                    float4 a3 = *reinterpret_cast<const float4 *>(inb + (size_t)j3 * C + v * 4);
                    acc[0] = fmaf(w0, a0.x, acc[0]); acc[1] = fmaf(w0, a0.y, acc[1]);
                    acc[2] = fmaf(w0, a0.z, acc[2]); acc[3] = fmaf(w0, a0.w, acc[3]);
                    acc[0] = fmaf(w1, a1.x, acc[0]); acc[1] = fmaf(w1, a1.y, acc[1]);
                    acc[2] = fmaf(w1, a1.z, acc[2]); acc[3] = fmaf(w1, a1.w, acc[3]);
                    acc[0] = fmaf(w2, a2.x, acc[0]); acc[1] = fmaf(w2, a2.y, acc[1]);
                    acc[2] = fmaf(w2, a2.z, acc[2]); acc[3] = fmaf(w2, a2.w, acc[3]);
                    acc[0] = fmaf(w3, a3.x, acc[0]); acc[1] = fmaf(w3, a3.y, acc[1]);
                    acc[2] = fmaf(w3, a3.z, acc[2]); acc[3] = fmaf(w3, a3.w, acc[3]);
                } else {
                    float a0 = inb[(size_t)j0 * C + v], a1 = inb[(size_t)j1 * C + v];
                    float a2 = inb[(size_t)j2 * C + v], a3 = inb[(size_t)j3 * C + v];
                    acc[0] = fmaf(w0, a0, acc[0]); acc[0] = fmaf(w1, a1, acc[0]);
                    acc[0] = fmaf(w2, a2, acc[0]); acc[0] = fmaf(w3, a3, acc[0]);
                }
            }
        }
        for (; k < cnt; ++k) {
            int j = __shfl_sync(0xffffffffu, my_col, k);
            float w = __shfl_sync(0xffffffffu, my_val, k);
            if (active) {
                if (VEC) {
                    float4 a = *reinterpret_cast<const float4 *>(inb + (size_t)j * C + v * 4);
                    acc[0] = fmaf(w, a.x, acc[0]); acc[1] = fmaf(w, a.y, acc[1]);
                    acc[2] = fmaf(w, a.z, acc[2]); acc[3] = fmaf(w, a.w, acc[3]);
                } else {
                    acc[0] = fmaf(w, inb[(size_t)j * C + v], acc[0]);
                }
            }
        }
    }
}

// One warp per (b, i) row.  VEC: lanes own float4 channel groups; else lanes own single channels.
// Rows with degree > HUB_DEG are skipped here when hubs are handled by gcn_aggregate_hub_kernel.
template <bool VEC>
__device__ __forceinline__ void hub_cta(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                        const float *__restrict__ val, const int32_t *__restrict__ hubs, int n_hubs,
                                        int Nv, const float *__restrict__ in, int C, int L,
                                        const float *__restrict__ bias, int relu, float *__restrict__ out,
                                        unsigned hub_block);

// Hub CTAs and row CTAs share one launch; hub CTAs are interleaved at a regular stride (block b is a hub
// CTA when b % stride == 0) so that their long neighbour lists start early, overlap with the short rows
// and never monopolise the SMs.
template <bool VEC>
__global__ void __launch_bounds__(AG_THREADS)
gcn_aggregate_kernel(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                     const float *__restrict__ val, const int32_t *__restrict__ hubs, int n_hubs, unsigned hub_ctas,
                     unsigned hub_stride, int Nv, const float *__restrict__ in, long long rows, int C, int L,
                     const float *__restrict__ bias, int relu, float *__restrict__ out) {
    unsigned hubs_before = 0;  // hub CTAs with a block id below this one
    if (hub_ctas > 0) {
        const unsigned slot = blockIdx.x / hub_stride;
        if (blockIdx.x % hub_stride == 0 && slot < hub_ctas) {
            hub_cta<VEC>(rowptr, col, val, hubs, n_hubs, Nv, in, C, L, bias, relu, out, slot);
            return;
        }
        hubs_before = slot + 1 < hub_ctas ? slot + 1 : hub_ctas;
    }
    const int skip_hubs = hub_ctas > 0;
    const long long row = (long long)(blockIdx.x - hubs_before) * AG_WARPS + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const int i = (int)(row % Nv);
    const long long b = row / Nv;
    const int beg = rowptr[i], end = rowptr[i + 1];
    const bool hub = skip_hubs && (end - beg) > HUB_DEG;
    const float *__restrict__ inb = in + (size_t)b * Nv * C;
    const float *__restrict__ self = in + (size_t)row * C;
    float *__restrict__ o = out + (size_t)row * C;
    const int W = VEC ? 4 : 1;
    const int ngroups = C / W;          // channel groups per row
    const int gath = (L + W - 1) / W;   // groups that contain at least one aggregated channel

    // pass-through groups (c >= L entirely)
    for (int v = gath + lane; v < ngroups; v += 32) {
        if (VEC) {
            float4 s = *reinterpret_cast<const float4 *>(self + v * 4);
            if (relu) {
                s.x = fmaxf(s.x, 0.f); s.y = fmaxf(s.y, 0.f); s.z = fmaxf(s.z, 0.f); s.w = fmaxf(s.w, 0.f);
            }
            *reinterpret_cast<float4 *>(o + v * 4) = s;
        } else {
            float s = self[v];
            o[v] = relu ? fmaxf(s, 0.f) : s;
        }
    }
    if (hub) return;
    // aggregated groups
    for (int v0 = 0; v0 < gath; v0 += 32) {
        const int v = v0 + lane;
        const bool active = v < gath;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        row_gather<VEC>(col, val, beg, end, inb, C, v, active, acc);
        if (!active) continue;
        if (VEC) {
            float4 s = *reinterpret_cast<const float4 *>(self + v * 4);
            float sv[4] = {s.x, s.y, s.z, s.w};
            float r[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c = v * 4 + k;
                float t = c < L ? acc[k] + (bias ? bias[c] : 0.f) : sv[k];
                r[k] = relu ? fmaxf(t, 0.f) : t;
            }
            *reinterpret_cast<float4 *>(o + v * 4) = make_float4(r[0], r[1], r[2], r[3]);
        } else {
            float t = acc[0] + (bias ? bias[v] : 0.f);
            o[v] = relu ? fmaxf(t, 0.f) : t;
        }
    }
}

// One CTA per (b, hub row): the 8 warps split the neighbour list; shared-memory reduction.
// Only the aggregated channel groups are written (pass-through was done by the row kernel).
template <bool VEC>
__device__ __forceinline__ void hub_cta(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                        const float *__restrict__ val, const int32_t *__restrict__ hubs, int n_hubs,
                                        int Nv, const float *__restrict__ in, int C, int L,
                                        const float *__restrict__ bias, int relu, float *__restrict__ out,
                                        unsigned hub_block) {
    const int i = hubs[hub_block % n_hubs];
    const long long b = hub_block / n_hubs;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int beg = rowptr[i], end = rowptr[i + 1];
    const int per = (((end - beg) + AG_WARPS - 1) / AG_WARPS + 31) / 32 * 32;
    const int wbeg = min(end, beg + warp * per), wend = min(end, wbeg + per);
    const float *__restrict__ inb = in + (size_t)b * Nv * C;
    const long long row = b * Nv + i;
    const float *__restrict__ self = in + (size_t)row * C;
    float *__restrict__ o = out + (size_t)row * C;
    const int W = VEC ? 4 : 1;
    const int gath = (L + W - 1) / W;
    __shared__ float s_part[AG_WARPS][32][4];
    for (int v0 = 0; v0 < gath; v0 += 32) {
        const int v = v0 + lane;
        const bool active = v < gath;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        row_gather<VEC>(col, val, wbeg, wend, inb, C, v, active, acc);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; ++k) s_part[warp][lane][k] = acc[k];
        __syncthreads();
        if (warp == 0 && active) {
            float tot[4] = {0.f, 0.f, 0.f, 0.f};
            for (int w = 0; w < AG_WARPS; ++w)
#pragma unroll
                for (int k = 0; k < 4; ++k) tot[k] += s_part[w][lane][k];
            if (VEC) {
                float4 s = *reinterpret_cast<const float4 *>(self + v * 4);
                float sv[4] = {s.x, s.y, s.z, s.w};
                float r[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int c = v * 4 + k;
                    float t = c < L ? tot[k] + (bias ? bias[c] : 0.f) : sv[k];
                    r[k] = relu ? fmaxf(t, 0.f) : t;
                }
                *reinterpret_cast<float4 *>(o + v * 4) = make_float4(r[0], r[1], r[2], r[3]);
            } else {
                float t = tot[0] + (bias ? bias[v] : 0.f);
                o[v] = relu ? fmaxf(t, 0.f) : t;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Tile kernel (the fast path: C % 4 == 0, at most 96 aggregated float4 groups).
//
// ncu on the one-warp-per-row kernel above (B=256, N=1949, C=300, L=99): DRAM traffic equals the
// algorithmic bytes, but only 35 % of the HBM rate is reached -- 511 warp instructions per row, mostly
// index arithmetic, shuffles and 64-bit divisions, and a serial rowptr -> col -> neighbour load chain.
// This kernel removes both:
//   * a CTA owns TV consecutive vertices and BG consecutive batch elements.  The graph is the same for
//     every batch element, so each warp stages the (byte offset, weight) list of its row ONCE in its
//     private shared-memory strip and replays it for the BG batch elements with 128-bit broadcast loads;
//   * per neighbour: 1/4 LDS.128 x2, one address add, one LDG.128, four FFMA; eight neighbour rows in
//     flight per lane, the pass-through part of the row is requested before the gather starts;
//   * the 8 warps of a CTA work on neighbouring rows of the same batch element at the same time, so the
//     neighbour rows they share (19-vertex charts) are served by L1;
//   * rows of degree > HUB_DEG are left to dedicated hub CTAs at the front of the grid (8 warps split the
//     neighbour list, fixed-order shared-memory reduction => deterministic).

template <int NG>
__global__ void __launch_bounds__(AG_THREADS, 4)
gcn_aggregate_tile_kernel(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                          const float *__restrict__ val, const AggHubs hb, unsigned hub_slots, int Nv,
                          const float *__restrict__ in, int B, int C, int L, const float *__restrict__ bias,
                          int relu, float *__restrict__ out, int TV, int BG, int n_tiles, int hubs_first, int ldi,
                          int ldo, int warp_is_batch, int prefetch_next) {
    // ldi / ldo: row strides (floats) of in / out; C channels are handled ([0, L) aggregated, [L, C) passed through)
    __shared__ __align__(16) uint32_t s_off[AG_WARPS][AT_STRIP];
    __shared__ __align__(16) float s_w[AG_WARPS][AT_STRIP];
    __shared__ __align__(16) float s_part[AG_WARPS][NG * 32 * 4];
    pdl_launch_dependents();
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ngroups = C >> 2;
    const int gath = (L + 3) >> 2;
    const uint32_t row_bytes = (uint32_t)ldi * 4u;
    bool on[NG];
    uint32_t voff[NG];  // byte offset of the lane's channel group (clamped to the last aggregated group)
#pragma unroll
    for (int n = 0; n < NG; ++n) {
        on[n] = lane + 32 * n < gath;
        voff[n] = (uint32_t)min(lane + 32 * n, gath - 1) * 16u;
    }
    const size_t bstride = (size_t)Nv * ldi, bstride_o = (size_t)Nv * ldo;  // floats per batch element
    const int npass = ngroups - gath;       // pure pass-through groups
    const bool pass0 = lane < npass, pass1 = 32 + lane < npass;
    const int pv0 = (gath + lane) * 4, pv1 = (gath + 32 + lane) * 4;  // float offsets inside the row

    // One output row of one batch element by one warp.  `acc` arrives initialised (0, or alpha * y*);
    // the row's (offset, weight) list is in the warp's strip when staged, else it is streamed in chunks.
    auto do_row = [&](int i, const float *inb, float *outb, int beg, int end, int n4, float (&acc)[NG][4],
                      int strip) {
        const float *self = inb + (size_t)i * ldi;
        float *o = outb + (size_t)i * ldo;
        // request the pass-through part of the row first (up to two groups per lane stay in registers)
        float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f), p1 = p0;
        if (pass0) p0 = __ldcs(reinterpret_cast<const float4 *>(self + pv0));
        if (pass1) p1 = __ldcs(reinterpret_cast<const float4 *>(self + pv1));
        const char *base[NG];
#pragma unroll
        for (int n = 0; n < NG; ++n) base[n] = reinterpret_cast<const char *>(inb) + voff[n];
        if (end - beg <= AT_STRIP) {
            strip_gather<NG>(s_off[strip], s_w[strip], n4, base, acc);
        } else {  // long row: re-stage chunk by chunk (own strip)
            for (int e0 = beg; e0 < end; e0 += AT_STRIP) {
                const int m4 = strip_stage(col, val, e0, min(AT_STRIP, end - e0), row_bytes, s_off[warp], s_w[warp]);
                strip_gather<NG>(s_off[warp], s_w[warp], m4, base, acc);
            }
        }
        row_epilogue<NG>(acc, on, self, o, bias, L, relu);
        if (pass0) __stcs(reinterpret_cast<float4 *>(o + pv0), relu4(p0, relu));
        if (pass1) __stcs(reinterpret_cast<float4 *>(o + pv1), relu4(p1, relu));
        for (int v = gath + 64 + lane; v < ngroups; v += 32)
            __stcs(reinterpret_cast<float4 *>(o + v * 4), relu4(__ldcs(reinterpret_cast<const float4 *>(self + v * 4)), relu));
    };

    // Block order: per batch group (BG batch elements) first its hub CTAs, then its tiles -- the boundary rows
    // the hub CTAs pull into L2 are the ones the group's tiles gather next.
    // (hubs_first: all hub CTAs lead the grid instead -- better when the whole input is L2-resident and the
    // only concern is starting the long hub rows early.)
    // (A persistent-CTA variant -- a few CTAs per SM walking the work items -- was measured slower at every batch
    // size: 314 vs 278 us at B=256, no gain at B=16.)
    const unsigned per_group = hub_slots + (unsigned)n_tiles;
    unsigned group, local;
    if (hubs_first) {
        const unsigned n_groups = ((unsigned)B + BG - 1) / BG;
        const unsigned lead = hub_slots * n_groups;
        if (blockIdx.x < lead) {
            group = blockIdx.x / hub_slots;
            local = blockIdx.x % hub_slots;
        } else {
            group = (blockIdx.x - lead) / (unsigned)n_tiles;
            local = hub_slots + (blockIdx.x - lead) % (unsigned)n_tiles;
        }
    } else {
        group = blockIdx.x / per_group;
        local = blockIdx.x % per_group;
    }
    const unsigned hub_ctas = hub_slots;  // > 0: rows above HUB_DEG / flagged rows belong to hub CTAs
    if (local < hub_slots) {
        float acc[NG][4];
#pragma unroll
        for (int n = 0; n < NG; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
        const bool common = hb.n_common > 0;
        // common mode: CTA = batch element, the warps split the common set;
        // plain mode:  CTA = (batch element, hub row), the warps split that row's neighbour list
        const int b = (int)group * BG + (common ? (int)local : (int)(local / hb.n_hubs));
        if (b >= B) return;
        const int hrow = common ? -1 : hb.hubs[local % hb.n_hubs];
        const int32_t *lcol = common ? hb.common_col : col;
        const float *lval = common ? hb.common_w : val;
        const int beg = common ? 0 : rowptr[hrow], end = common ? hb.n_common : rowptr[hrow + 1];
        const int per = (((end - beg) + AG_WARPS - 1) / AG_WARPS + 3) & ~3;
        const int wbeg = min(end, beg + warp * per), wend = min(end, wbeg + per);
        const float *inb = in + b * bstride;
        float *outb = out + b * bstride_o;
        const char *base[NG];
#pragma unroll
        for (int n = 0; n < NG; ++n) base[n] = reinterpret_cast<const char *>(inb) + voff[n];
        for (int e0 = wbeg; e0 < wend; e0 += AT_STRIP) {
            const int n4 = strip_stage(lcol, lval, e0, min(AT_STRIP, wend - e0), row_bytes, s_off[warp], s_w[warp]);
            strip_gather<NG>(s_off[warp], s_w[warp], n4, base, acc);
        }
#pragma unroll
        for (int n = 0; n < NG; ++n)
#pragma unroll
            for (int k = 0; k < 4; ++k) s_part[warp][(n * 32 + lane) * 4 + k] = acc[n][k];
        __syncthreads();
        // fixed-order reduction over the warps (every warp computes the same total: deterministic)
#pragma unroll
        for (int n = 0; n < NG; ++n)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float t = 0.f;
                for (int w = 0; w < AG_WARPS; ++w) t += s_part[w][(n * 32 + lane) * 4 + k];
                acc[n][k] = t;
            }
        if (!common) {
            if (warp == 0) {
                row_epilogue<NG>(acc, on, inb + (size_t)hrow * ldi, outb + (size_t)hrow * ldo, bias, L, relu);
            } else {
                const float *self = inb + (size_t)hrow * ldi;
                float *o = outb + (size_t)hrow * ldo;
                for (int v = gath + (warp - 1) * 32 + lane; v < ngroups; v += (AG_WARPS - 1) * 32)
                    __stcs(reinterpret_cast<float4 *>(o + v * 4), relu4(__ldcs(reinterpret_cast<const float4 *>(self + v * 4)), relu));
            }
            return;
        }
        // acc = y*.  Each warp finishes some of the hub rows: alpha * y* + the reduced row.
        for (int h = warp; h < hb.n_hubs; h += AG_WARPS) {
            const int i = hb.hubs[h];
            const float a = hb.alpha[h];
            const int rb = rowptr[i], re = rowptr[i + 1];
            int n4 = 0;
            if (re - rb <= AT_STRIP) n4 = strip_stage(col, val, rb, re - rb, row_bytes, s_off[warp], s_w[warp]);
            float r[NG][4];
#pragma unroll
            for (int n = 0; n < NG; ++n)
#pragma unroll
                for (int k = 0; k < 4; ++k) r[n][k] = a * acc[n][k];
            do_row(i, inb, outb, rb, re, n4, r, warp);
        }
        return;
    }

    const int i0 = (int)(local - hub_slots) * TV;
    const int b0 = (int)group * BG;
    const int nb = min(B, b0 + BG) - b0;
    const int i1 = min(Nv, i0 + TV);
    if (warp_is_batch && TV == AG_WARPS && BG == AG_WARPS) {
        // ---- "warp = batch element" mapping: the 8 strips of the tile are staged once by the CTA; warp w then walks the
        // tile's 8 rows for batch element b0 + w.  Consecutive rows of a chart share most of their neighbours, so a
        // warp re-touches the same few input rows back to back and finds them in L1 -- without any CTA barrier
        // beyond the one after staging.
        __shared__ int s_meta[AG_WARPS][3];  // beg, end, staged count (-1: row not handled here)
        {
            const int i = i0 + warp;
            int beg = 0, end = 0, n4 = -1;
            if (i < i1) {
                beg = rowptr[i];
                end = rowptr[i + 1];
                const bool skip = hb.row_skip ? hb.row_skip[i] != 0 : (hub_ctas > 0 && end - beg > HUB_DEG);
                if (!skip) n4 = end - beg <= AT_STRIP ? strip_stage(col, val, beg, end - beg, row_bytes, s_off[warp], s_w[warp]) : -2;
            }
            if (lane == 0) {
                s_meta[warp][0] = beg;
                s_meta[warp][1] = end;
                s_meta[warp][2] = n4;
            }
        }
        const int any_long = __syncthreads_or(0);  // (also the barrier that publishes strips and meta)
        (void)any_long;
        bool has_long = false;
#pragma unroll
        for (int r = 0; r < AG_WARPS; ++r) has_long |= s_meta[r][2] == -2;
        if (!has_long) {
            const int b = b0 + warp;
            if (b >= B) return;
            const float *inb = in + (size_t)b * bstride;
            float *outb = out + (size_t)b * bstride_o;
            for (int r = 0; r < AG_WARPS; ++r) {
                const int n4 = s_meta[r][2];
                if (n4 < 0) continue;
                float acc[NG][4];
#pragma unroll
                for (int n = 0; n < NG; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
                do_row(i0 + r, inb, outb, s_meta[r][0], s_meta[r][1], n4, acc, r);
            }
            return;
        }
        __syncthreads();  // a row longer than the strip: fall through to the row-per-warp mapping (re-stages)
    }
    for (int i = i0 + warp; i < i1; i += AG_WARPS) {
        const int beg = rowptr[i], end = rowptr[i + 1];
        if (hb.row_skip ? hb.row_skip[i] != 0 : (hub_ctas > 0 && end - beg > HUB_DEG)) continue;  // hub CTA's row
        int n4 = 0;
        if (end - beg <= AT_STRIP) n4 = strip_stage(col, val, beg, end - beg, row_bytes, s_off[warp], s_w[warp]);
        const float *inb = in + b0 * bstride;
        float *outb = out + b0 * bstride_o;
        const int row_lines = (int)((ngroups * 16 + 127) >> 7);  // 128-byte lines of one input row
        for (int bb = 0; bb < nb; ++bb, inb += bstride, outb += bstride_o) {
            // The next batch element's copy of this row is its first touch (a DRAM miss): ask L2 for it now, one
            // row (~2.5 us of this warp's work) ahead -- no registers held, unlike a software-pipelined load.
            // Every row of `in` is requested this way by the warp that owns it, so the neighbour gathers of the
            // other warps find their rows in L2 as well.
            if (prefetch_next && bb + 1 < nb && lane < row_lines)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(inb + bstride + (size_t)i * ldi) + lane * 128));
            float acc[NG][4];
#pragma unroll
            for (int n = 0; n < NG; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
            do_row(i, inb, outb, beg, end, n4, acc, warp);
        }
    }
}

// Narrow rows (C <= 8, e.g. the 3-channel last layer): one thread per output row, hub rows by one warp each.
template <int CM>
__global__ void __launch_bounds__(256)
gcn_aggregate_narrow_kernel(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                            const float *__restrict__ val, const int32_t *__restrict__ hubs, int n_hubs,
                            unsigned hub_ctas, int Nv, const float *__restrict__ in, long long rows, int B, int C, int L,
                            const float *__restrict__ bias, int relu, float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    if (blockIdx.x < hub_ctas) {
        const long long hw = (long long)blockIdx.x * AG_WARPS + (threadIdx.x >> 5);  // (b, hub) pairs
        if (hw >= (long long)B * n_hubs) return;
        const int i = hubs[hw % n_hubs];
        const long long b = hw / n_hubs;
        const float *inb = in + (size_t)b * Nv * C;
        float acc[CM];
#pragma unroll
        for (int c = 0; c < CM; ++c) acc[c] = 0.f;
        for (int e = rowptr[i] + lane; e < rowptr[i + 1]; e += 32) {
            const float w = val[e];
            const float *src = inb + (size_t)col[e] * C;
#pragma unroll
            for (int c = 0; c < CM; ++c)
                if (c < L) acc[c] = fmaf(w, src[c], acc[c]);
        }
#pragma unroll
        for (int c = 0; c < CM; ++c) acc[c] = warp_sum(acc[c]);
        if (lane == 0) {
            const size_t row = (size_t)b * Nv + i;
#pragma unroll
            for (int c = 0; c < CM; ++c)
                if (c < C) {
                    float t = c < L ? acc[c] + (bias ? bias[c] : 0.f) : in[row * C + c];
                    out[row * C + c] = relu ? fmaxf(t, 0.f) : t;
                }
        }
        return;
    }
    const long long row = (long long)(blockIdx.x - hub_ctas) * 256 + threadIdx.x;
    if (row >= rows) return;
    const int i = (int)(row % Nv);
    const float *inb = in + (size_t)(row - i) * C;
    const int beg = rowptr[i], end = rowptr[i + 1];
    if (hub_ctas > 0 && end - beg > HUB_DEG) return;
    float acc[CM];
#pragma unroll
    for (int c = 0; c < CM; ++c) acc[c] = 0.f;
    for (int e = beg; e < end; ++e) {
        const float w = val[e];
        const float *src = inb + (size_t)col[e] * C;
#pragma unroll
        for (int c = 0; c < CM; ++c)
            if (c < L) acc[c] = fmaf(w, src[c], acc[c]);
    }
#pragma unroll
    for (int c = 0; c < CM; ++c)
        if (c < C) {
            float t = c < L ? acc[c] + (bias ? bias[c] : 0.f) : in[(size_t)row * C + c];
            out[(size_t)row * C + c] = relu ? fmaxf(t, 0.f) : t;
        }
}

// gbias[c] = sum_rows g[row, c] (c < L), 0 otherwise.  Deterministic two-stage column sum:
// stage 1: CTA `k` sums rows k, k+G, ... into part[k, c]; stage 2: one CTA sums the G partials.
constexpr int BG_PARTS = 1184;  // 8 slabs per SM: short dependent-load chains in stage 1
// CTA k sums its contiguous slab of rows; 8 row lanes x 32 column lanes (128-byte coalesced segments),
// fixed-order shared-memory reduction over the row lanes.
__global__ void __launch_bounds__(256)
bias_grad_stage1(const float *__restrict__ g, long long M, int C, int L, float *__restrict__ part) {
    pdl_launch_dependents();
    pdl_wait();
    // blockIdx.y = layer of a batched call: g and part advance by one (M x C) matrix / one partial block per layer
    g += (size_t)blockIdx.y * (size_t)M * C;
    part += (size_t)blockIdx.y * gridDim.x * L;
    // four 32-column blocks per pass: 4 independent loads per row and thread in flight, one barrier pair per pass
    constexpr int NB = 4;
    __shared__ float red[NB][8][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const long long per = (M + gridDim.x - 1) / gridDim.x;
    const long long r0 = (long long)blockIdx.x * per;
    const long long r1 = r0 + per < M ? r0 + per : M;
    for (int cbase = 0; cbase < L; cbase += 32 * NB) {
        float acc[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) acc[j] = 0.f;
#pragma unroll 4
        for (long long r = r0 + ry; r < r1; r += 8) {
            const float *row = g + (size_t)r * C + cbase + cx;
#pragma unroll
            for (int j = 0; j < NB; ++j)
                if (cbase + 32 * j + cx < L) acc[j] += row[32 * j];
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) red[j][ry][cx] = acc[j];
        __syncthreads();
        if (ry < NB) {  // warp j finishes column block j (fixed order over the row lanes)
            const int c = cbase + 32 * ry + cx;
            if (c < L) {
                float t = red[ry][0][cx];
#pragma unroll
                for (int y = 1; y < 8; ++y) t += red[ry][y][cx];
                part[(size_t)blockIdx.x * L + c] = t;
            }
        }
        __syncthreads();
    }
}
// stage 2: one warp per column; lanes stride over the partials, fixed-order butterfly => deterministic
__global__ void __launch_bounds__(256)
bias_grad_stage2(const float *__restrict__ part, int nparts, int C, int L, float *__restrict__ gbias) {
    pdl_launch_dependents();
    pdl_wait();
    part += (size_t)blockIdx.y * nparts * L;
    gbias += (size_t)blockIdx.y * C;
    const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (c >= C) return;
    float acc = 0.f;
    if (c < L)
        for (int p = lane; p < nparts; p += 32) acc += part[(size_t)p * L + c];
    acc = warp_sum(acc);
    if (lane == 0) gbias[c] = acc;
}

// gpre = (act > 0) ? g : 0  -- ReLU backward for the stand-alone layer / last-layer cases
__global__ void relu_mask_kernel(const float *__restrict__ g, const float *__restrict__ act, long long n,
                                 float *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = act[i] > 0.f ? g[i] : 0.f;
}

}  // namespace ptk

using namespace ptk;

extern "C" int ptk_relu_mask(const float *g, const float *act, int64_t n, float *out, ptk_stream_t stream) {
    PTK_REQUIRE(g && act && out && n >= 0, PTK_ERR_SHAPE, "relu_mask: bad arguments");
    if (n == 0) return PTK_OK;
    relu_mask_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(g, act, (long long)n, out);
    PTK_CHECK_LAUNCH();
    return PTK_OK;
}

extern "C" int ptk_gcn_aggregate_ex(const int32_t *rowptr, const int32_t *col, const float *val,
                                    const int32_t *hubs, int32_t n_hubs, const int32_t *common_col,
                                    const float *common_w, int32_t n_common, const float *hub_alpha,
                                    const uint8_t *row_skip, int64_t Nv, const float *in, int64_t B,
                                    int64_t C, int64_t L, const float *bias, int relu, float *out,
                                    int64_t ldi, int64_t ldo, ptk_stream_t stream) {
    PTK_NVTX("ptk_gcn_aggregate_ex");
    if (ldi <= 0) ldi = C;
    if (ldo <= 0) ldo = C;
    PTK_REQUIRE(rowptr && col && val && in && out, PTK_ERR_SHAPE, "gcn_aggregate: null pointer");
    PTK_REQUIRE(B > 0 && Nv > 0 && C > 0 && L >= 0 && L <= C, PTK_ERR_SHAPE,
                "gcn_aggregate: bad sizes (B=%lld, Nv=%lld, C=%lld, L=%lld)", (long long)B,
                (long long)Nv, (long long)C, (long long)L);
    PTK_REQUIRE(in != out, PTK_ERR_SHAPE, "gcn_aggregate: in-place aggregation is not supported");
    PTK_REQUIRE(ldi >= C && ldo >= C, PTK_ERR_SHAPE, "gcn_aggregate: row strides must be >= C");
    const bool strided = ldi != C || ldo != C;
    const bool common = n_common > 0;
    PTK_REQUIRE(!common || (hubs && n_hubs > 0 && common_col && common_w && hub_alpha && row_skip), PTK_ERR_SHAPE,
                "gcn_aggregate: a common neighbour set needs hubs, common_col, common_w, hub_alpha and row_skip");
    cudaStream_t st = as_stream(stream);
    const long long rows = (long long)B * Nv;
    const bool vec = (C % 4 == 0) && (ldi % 4 == 0) && (ldo % 4 == 0) &&
                     ((((uintptr_t)in) | ((uintptr_t)out) | ((uintptr_t)bias)) % 16 == 0);
    const bool have_hubs = hubs && n_hubs > 0;
    const int gath = (int)((L + 3) / 4);
    PTK_REQUIRE(B <= 0x7fffffff && Nv * ldi * 4 <= 0xffffffffLL, PTK_ERR_SHAPE, "gcn_aggregate: one batch element must be < 4 GiB");
    if (vec && gath >= 1 && gath <= 96) {
        // TV = 8 (one row per warp) keeps few batch elements in flight at a time: the rows a batch element
        // gathers stay in L2 until all its tiles are done.  BG amortises the staging of the row structure.
        int TV = 8, BG = 8;
        // >= 6 CTAs per SM before BG is cut: at B = 16 (config 3) BG = 4 -> 976 CTAs measured 10 % faster than the
        // 3904 single-element CTAs a larger target gives (tools/agg_bench.py 16, PTK_AGG_BG sweep)
        const long long slots = 6LL * sm_count();
        while (BG > 1 && ceil_div(Nv, TV) * ceil_div(B, BG) < slots) BG >>= 1;
        if (PTK_TUNING_ENV("PTK_AGG_TV") > 0) TV = PTK_TUNING_ENV("PTK_AGG_TV");  // tools/agg_bench.py sweeps
        if (PTK_TUNING_ENV("PTK_AGG_BG") > 0) BG = PTK_TUNING_ENV("PTK_AGG_BG");
        const int n_tiles = (int)ceil_div(Nv, TV);
        const unsigned hub_slots = have_hubs ? (unsigned)(common ? BG : BG * n_hubs) : 0u;  // per batch group
        const unsigned grid = (unsigned)((hub_slots + n_tiles) * ceil_div(B, BG));
        const int hubs_first = (double)B * Nv * (ldi + ldo) * 4.0 < 100e6;  // input + output fit the 126 MB L2
        // measured at B=256: +8 % when more than 128 channels are aggregated (L=300), +1..2 % on the fused touch graphs
        // and -3 % on the plain vision graph at L=99 -> used for the wide case only
        int warp_is_batch = TV == AG_WARPS && BG == AG_WARPS && gath > 32;
        if (PTK_TUNING_ENV("PTK_AGG_WB") > 0) warp_is_batch = PTK_TUNING_ENV("PTK_AGG_WB") == 1;
        // L2 prefetch of the next batch element's row: only where the input does not live in L2 anyway
        int prefetch_next = !hubs_first;
        if (PTK_TUNING_ENV("PTK_AGG_PF") > 0) prefetch_next = PTK_TUNING_ENV("PTK_AGG_PF") == 1;
        AggHubs hb;
        hb.hubs = hubs; hb.n_hubs = have_hubs ? n_hubs : 0;
        hb.common_col = common_col; hb.common_w = common_w; hb.n_common = common ? n_common : 0;
        hb.alpha = hub_alpha; hb.row_skip = common ? row_skip : nullptr;
#define PTK_TILE(NGv)                                                                                        \
    launch_pdl(gcn_aggregate_tile_kernel<NGv>, dim3(grid), dim3(AG_THREADS), 0, st, rowptr, col, val, hb, hub_slots, (int)Nv, in, \
               (int)B, (int)C, (int)L, bias, relu, out, TV, BG, n_tiles, hubs_first, (int)ldi, (int)ldo, warp_is_batch, \
               prefetch_next)
        if (gath <= 32) PTK_TILE(1);
        else if (gath <= 64) PTK_TILE(2);
        else PTK_TILE(3);
#undef PTK_TILE
        PTK_CHECK_LAUNCH();
        return PTK_OK;
    }
    PTK_REQUIRE(!strided, PTK_ERR_SHAPE, "gcn_aggregate: row strides need the vector path (C, ldi, ldo %% 4 == 0, 1 <= L <= 384)");
    PTK_REQUIRE(!common, PTK_ERR_SHAPE,
                "gcn_aggregate: the common-set form needs the vector path (C %% 4 == 0, 1 <= L, ceil(L/4) <= 96, "
                "16-byte aligned); pass the unreduced CSR otherwise");
    if (C <= 8) {
        const unsigned hub_ctas = have_hubs ? (unsigned)ceil_div(B * n_hubs, AG_WARPS) : 0u;
        const unsigned grid = hub_ctas + (unsigned)ceil_div(rows, 256);
        if (C <= 4)
            gcn_aggregate_narrow_kernel<4><<<grid, 256, 0, st>>>(rowptr, col, val, hubs, n_hubs, hub_ctas, (int)Nv, in, rows, (int)B, (int)C, (int)L, bias, relu, out);
        else
            gcn_aggregate_narrow_kernel<8><<<grid, 256, 0, st>>>(rowptr, col, val, hubs, n_hubs, hub_ctas, (int)Nv, in, rows, (int)B, (int)C, (int)L, bias, relu, out);
        PTK_CHECK_LAUNCH();
        return PTK_OK;
    }
    const bool vec_old = (C % 4 == 0) && ((((uintptr_t)in) | ((uintptr_t)out)) % 16 == 0);
    const unsigned hub_ctas = have_hubs ? (unsigned)(B * n_hubs) : 0u;
    const unsigned grid = hub_ctas + (unsigned)ceil_div(rows, AG_WARPS);
    const unsigned hub_stride = hub_ctas > 0 ? (grid / hub_ctas > 0 ? grid / hub_ctas : 1u) : 1u;
    if (vec_old)
        gcn_aggregate_kernel<true><<<grid, AG_THREADS, 0, st>>>(rowptr, col, val, hubs, n_hubs, hub_ctas, hub_stride, (int)Nv, in, rows, (int)C, (int)L, bias, relu, out);
    else
        gcn_aggregate_kernel<false><<<grid, AG_THREADS, 0, st>>>(rowptr, col, val, hubs, n_hubs, hub_ctas, hub_stride, (int)Nv, in, rows, (int)C, (int)L, bias, relu, out);
    PTK_CHECK_LAUNCH();
    return PTK_OK;
}

namespace ptk {
struct AggTiles {
    const int32_t *uptr;
    const int32_t *ucol;
    const uint16_t *lidx;
};
int aggregate_union_launch(int mode, const int32_t *rowptr, const int32_t *col, const float *val, const AggHubs &hb,
                           const AggTiles &tl, int max_union, bool have_hubs, int64_t Nv, const float *in, int64_t B, int64_t C,
                           int64_t L, const float *bias, int relu, float *out, int64_t ldi, int64_t ldo, int hubs_first,
                           int prefetch_next, cudaStream_t st);  // gcn_aggregate_union.cu
}  // namespace ptk

extern "C" int ptk_gcn_aggregate_tiled(const int32_t *rowptr, const int32_t *col, const float *val,
                                       const int32_t *hubs, int32_t n_hubs, const int32_t *common_col,
                                       const float *common_w, int32_t n_common, const float *hub_alpha,
                                       const uint8_t *row_skip, const int32_t *tile_uptr, const int32_t *tile_ucol,
                                       const uint16_t *tile_lidx, int32_t max_union, int32_t mode, int64_t Nv,
                                       const float *in, int64_t B, int64_t C, int64_t L, const float *bias, int relu,
                                       float *out, int64_t ldi, int64_t ldo, ptk_stream_t stream) {
    const int64_t ldi_e = ldi <= 0 ? C : ldi, ldo_e = ldo <= 0 ? C : ldo;
    const bool vec = rowptr && col && val && in && out && in != out && B > 0 && Nv > 0 && C > 0 && L >= 1 && L <= C &&
                     (C % 4 == 0) && (ldi_e % 4 == 0) && (ldo_e % 4 == 0) && ldi_e >= C && ldo_e >= C &&
                     ((((uintptr_t)in) | ((uintptr_t)out) | ((uintptr_t)bias)) % 16 == 0) && B <= 0x7fffffff &&
                     Nv * ldi_e * 4 <= 0xffffffffLL;
    const bool common = n_common > 0;
    const bool common_ok = !common || (hubs && n_hubs > 0 && common_col && common_w && hub_alpha && row_skip);
    PTK_REQUIRE(mode >= PTK_AGG_AUTO && mode <= PTK_AGG_RING, PTK_ERR_SHAPE, "gcn_aggregate_tiled: unknown mode %d", mode);
    // PTK_AGG_AUTO: the form measured fastest on B200 (profiles/r02_gcn_aggregate_forms.txt) -- the dense-tile product for
    // wide layers (more than 128 aggregated channels) and when nothing is passed through (the compact head of the fused
    // layer); the L2 gather where the same warp would also have to copy pass-through columns.
    int form = mode;
    if (mode == PTK_AGG_AUTO) {
        const bool wide = L > 128;
        const bool no_pass = (L + 3) / 4 == C / 4;
        form = (wide || no_pass) ? PTK_AGG_DENSE_TILE : PTK_AGG_L2_GATHER;
    }
    if (form != PTK_AGG_L2_GATHER && vec && common_ok && tile_uptr && tile_ucol && tile_lidx && max_union > 0) {
        AggHubs hb;
        const bool have_hubs = hubs && n_hubs > 0;
        hb.hubs = hubs; hb.n_hubs = have_hubs ? n_hubs : 0;
        hb.common_col = common_col; hb.common_w = common_w; hb.n_common = common ? n_common : 0;
        hb.alpha = hub_alpha; hb.row_skip = common ? row_skip : nullptr;
        AggTiles tl;
        tl.uptr = tile_uptr; tl.ucol = tile_ucol; tl.lidx = tile_lidx;
        const int hubs_first = (double)B * Nv * (ldi_e + ldo_e) * 4.0 < 100e6;
        int prefetch_next = !hubs_first;
        if (PTK_TUNING_ENV("PTK_AGG_PF") > 0) prefetch_next = PTK_TUNING_ENV("PTK_AGG_PF") == 1;
        const int rc = aggregate_union_launch(form == PTK_AGG_DENSE_TILE ? 1 : 2, rowptr, col, val, hb, tl, max_union, have_hubs, Nv,
                                              in, B, C, L, bias, relu, out, ldi_e, ldo_e, hubs_first, prefetch_next,
                                              as_stream(stream));
        if (rc != 0) return rc < 0 ? rc : PTK_OK;
    }
    return ptk_gcn_aggregate_ex(rowptr, col, val, hubs, n_hubs, common_col, common_w, n_common, hub_alpha, row_skip, Nv, in,
                                B, C, L, bias, relu, out, ldi, ldo, stream);
}

extern "C" int ptk_gcn_aggregate(const int32_t *rowptr, const int32_t *col, const float *val,
                                 const int32_t *hubs, int32_t n_hubs, int64_t Nv, const float *in,
                                 int64_t B, int64_t C, int64_t L, const float *bias, int relu,
                                 float *out, ptk_stream_t stream) {
    return ptk_gcn_aggregate_ex(rowptr, col, val, hubs, n_hubs, nullptr, nullptr, 0, nullptr, nullptr, Nv, in, B,
                                C, L, bias, relu, out, 0, 0, stream);
}

extern "C" size_t ptk_gcn_bias_grad_workspace_bytes(int64_t M, int64_t L) {
    const int64_t nparts = M < BG_PARTS ? M : BG_PARTS;
    return sizeof(float) * (size_t)(nparts > 0 ? nparts : 1) * (size_t)(L > 0 ? L : 1);
}

extern "C" size_t ptk_gcn_bias_grad_batched_workspace_bytes(int64_t n_mats, int64_t M, int64_t L) {
    return (size_t)(n_mats > 0 ? n_mats : 1) * ptk_gcn_bias_grad_workspace_bytes(M, L);
}

extern "C" int ptk_gcn_bias_grad_batched(const float *g, int64_t n_mats, int64_t M, int64_t C, int64_t L,
                                         float *gbias, void *workspace, size_t workspace_bytes,
                                         ptk_stream_t stream) {
    PTK_REQUIRE(g && gbias, PTK_ERR_SHAPE, "gcn_bias_grad_batched: null pointer");
    PTK_REQUIRE(n_mats > 0 && n_mats <= 65535 && M > 0 && C > 0 && L >= 0 && L <= C, PTK_ERR_SHAPE,
                "gcn_bias_grad_batched: bad sizes");
    PTK_REQUIRE(workspace && workspace_bytes >= ptk_gcn_bias_grad_batched_workspace_bytes(n_mats, M, L),
                PTK_ERR_WORKSPACE, "gcn_bias_grad_batched: workspace too small");
    cudaStream_t st = as_stream(stream);
    const int nparts = (int)(M < BG_PARTS ? M : BG_PARTS);
    float *part = reinterpret_cast<float *>(workspace);
    if (L > 0) {
        launch_pdl(bias_grad_stage1, dim3((unsigned)nparts, (unsigned)n_mats), dim3(256), 0, st, g, (long long)M, (int)C, (int)L, part);
        PTK_CHECK_LAUNCH();
    }
    launch_pdl(bias_grad_stage2, dim3((unsigned)ceil_div(C, 8), (unsigned)n_mats), dim3(256), 0, st, part, nparts, (int)C, (int)L, gbias);
    PTK_CHECK_LAUNCH();
    return PTK_OK;
}

extern "C" int ptk_gcn_bias_grad(const float *g, int64_t M, int64_t C, int64_t L, float *gbias,
                                 void *workspace, size_t workspace_bytes, ptk_stream_t stream) {
    PTK_REQUIRE(g && gbias, PTK_ERR_SHAPE, "gcn_bias_grad: null pointer");
    PTK_REQUIRE(M > 0 && C > 0 && L >= 0 && L <= C, PTK_ERR_SHAPE, "gcn_bias_grad: bad sizes");
    PTK_REQUIRE(workspace && workspace_bytes >= ptk_gcn_bias_grad_workspace_bytes(M, L),
                PTK_ERR_WORKSPACE, "gcn_bias_grad: workspace too small");
    cudaStream_t st = as_stream(stream);
    const int nparts = (int)(M < BG_PARTS ? M : BG_PARTS);
    float *part = reinterpret_cast<float *>(workspace);
    if (L > 0) {
        launch_pdl(bias_grad_stage1, dim3((unsigned)nparts), dim3(256), 0, st, g, (long long)M, (int)C, (int)L, part);
        PTK_CHECK_LAUNCH();
    }
    launch_pdl(bias_grad_stage2, dim3((unsigned)ceil_div(C, 8)), dim3(256), 0, st, part, nparts, (int)C, (int)L, gbias);
    PTK_CHECK_LAUNCH();
    return PTK_OK;
}
