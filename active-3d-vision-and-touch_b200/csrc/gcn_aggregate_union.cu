// GCN vertex aggregation over per-tile neighbour unions (sm_100a): the dense-tile form and the shared-memory ring form.
//
// Same contract as gcn_aggregate_tile_kernel (gcn_aggregate.cu; replaces the dense `torch.matmul(adj, features[:, :, :L])`
// + cat + bias + activation of GCN_layer.forward, pterotactyl/reconstruction/vision/model.py:354-363):
//
//   out[b,i,c] = act( sum_{e in row i} val[e] * in[b,col[e],c] + bias[c] )   c <  L
//   out[b,i,c] = act( in[b,i,c] )                                             c >= L
//
// What the L2 gather leaves on the table (profiles/r01_ncu_gcn_aggregate_tile_v12.txt): DRAM traffic already equals
// the algorithmic bytes, but every neighbour row is fetched from L2 once per edge -- 9.5 (finger graph) to 16 (grasp
// graph) reads of each input row per batch element.  The 8 rows of a tile share most of their neighbours (19-vertex
// charts + their twins): the UNION of a tile's neighbour columns is 2.3-3.0x smaller than the sum of its degrees.  The
// host lists that union once per tile (graph.tile_unions); two kernels forms use it (mode of ptk_gcn_aggregate_tiled):
//
// DENSE TILE (PTK_AGG_DENSE_TILE) -- a warp owns the tile for one batch element and computes it as a small dense product
//   out[8 x C'] = A[8 x U] . X[U x C']: every union row is read from L2 once and accumulated into up to 8 register rows;
//   A (weights, 0 where a row does not use the column) sits in shared memory.  FMA work grows 1.6-3.5x, L2 reads fall
//   2.3-3.0x.  Measured (profiles/r02_gcn_aggregate_forms.txt, B = 256): L = C = 300 grasp graph 826 -> 648 us, finger graph
//   500 -> 467 us; compact head of the fused forward (no pass-through columns) 336 -> 267 us / 212 -> 195 us; SLOWER where
//   201 pass-through columns have to be copied by the same warp (finger graph 259 -> 294 us).  At the training batch
//   (B <= 32) it keeps 8 union rows in flight per lane at 2 CTAs per SM instead of 4 at 3 (17.7 -> 17.0 us on the finger
//   graph's compact head, 26.7 -> 21.3 us on the grasp graph's).  PTK_AGG_AUTO picks it where it wins: wide layers and
//   inputs without pass-through columns.
//
// RING (PTK_AGG_RING) -- per batch element the CTA brings the union's rows into shared memory with cp.async (every warp
//   copies its share of the rows, one row = one 16-byte LDGSTS per lane; completion is reported to the slot's mbarrier by
//   cp.async.mbarrier.arrive.noinc) into a 2-4 slot ring and gathers from shared memory (LDS.128 by local index) while the
//   copies of the next batch elements are in flight; no CTA barrier in the loop.  Wide layers in column chunks of <= 32
//   float4 groups.  MEASURED SLOWER than the L2 gather everywhere (profiles/r02_gcn_aggregate_union.txt): finger graph 348
//   vs 257 us, grasp graph 468 vs 393 us, B = 16: 29 vs 18 us.  Three producer schemes were tried: one cp.async.bulk per
//   row from warp 0 (458 us: the TMA unit retires a 400-byte bulk copy only every ~65 cycles per SM), a dedicated ninth
//   producer warp with cp.async (355 us) and the cooperative form kept here (348 us).  ncu: DRAM traffic again equals
//   the algorithmic bytes, long-scoreboard stalls fall from 7.3 to 3.6 per issue, but every gathered value crosses the
//   L1 / shared-memory pipe twice (LDGSTS in, LDS out) and the 8 warps of a tile move in lock step with their slot.
//   Kept for comparison; never picked by PTK_AGG_AUTO.
//
// Hub rows (degree > HUB_DEG) and their common neighbour set are handled exactly as in the tile kernel.
#include "gcn_aggregate_common.cuh"

namespace ptk {

constexpr int AU_TV = 8;          // rows per tile = consumer warps per CTA
constexpr int AU_THREADS = AG_THREADS;       // every warp both loads (its share of the union rows) and gathers
constexpr int AU_MAXU = 256;      // largest union (rows) a tile may have

__device__ __forceinline__ uint32_t au_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes (or the hint expires)
// instead of spinning -- without it the polling loop (TRYWAIT + YIELD + BRA) was 50 % of all issued instructions.
__device__ __forceinline__ void au_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(done) : "r"(bar), "r"(parity), "r"(1000000u) : "memory");
    } while (!done);
}
__device__ __forceinline__ float4 au_lds_f4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}

// gather of one row from a shared-memory stage: s_off holds LOCAL byte offsets (union index * stage row bytes)
__device__ __forceinline__ void strip_gather_smem(const uint32_t *__restrict__ s_off, const float *__restrict__ s_w, int n4,
                                                  uint32_t base, float (&acc)[4]) {
    int k = 0;
    for (; k + 8 <= n4; k += 8) {
        const uint4 o0 = *reinterpret_cast<const uint4 *>(s_off + k), o1 = *reinterpret_cast<const uint4 *>(s_off + k + 4);
        const float4 w0 = *reinterpret_cast<const float4 *>(s_w + k), w1 = *reinterpret_cast<const float4 *>(s_w + k + 4);
        const uint32_t off[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
        const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        float4 a[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) a[u] = au_lds_f4(base + off[u]);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            acc[0] = fmaf(w[u], a[u].x, acc[0]); acc[1] = fmaf(w[u], a[u].y, acc[1]);
            acc[2] = fmaf(w[u], a[u].z, acc[2]); acc[3] = fmaf(w[u], a[u].w, acc[3]);
        }
    }
    if (k < n4) {
        const uint4 o0 = *reinterpret_cast<const uint4 *>(s_off + k);
        const float4 w0 = *reinterpret_cast<const float4 *>(s_w + k);
        const uint32_t off[4] = {o0.x, o0.y, o0.z, o0.w};
        const float w[4] = {w0.x, w0.y, w0.z, w0.w};
        float4 a[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) a[u] = au_lds_f4(base + off[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            acc[0] = fmaf(w[u], a[u].x, acc[0]); acc[1] = fmaf(w[u], a[u].y, acc[1]);
            acc[2] = fmaf(w[u], a[u].z, acc[2]); acc[3] = fmaf(w[u], a[u].w, acc[3]);
        }
    }
}

struct AggTiles {
    const int32_t *uptr;   // (n_tiles + 1) offsets into ucol
    const int32_t *ucol;   // union columns of every tile, ascending within a tile
    const uint16_t *lidx;  // per CSR entry: index of col[e] in its row's tile union
};

// XW: union rows a lane of the dense-tile form keeps in flight -- 4 at three CTAs per SM (large batches: occupancy
// matters), 8 at two CTAs per SM (training batch: few CTAs per SM exist anyway, memory-level parallelism per warp matters)
template <int NG, int XW>
__global__ void __launch_bounds__(AU_THREADS, XW == 8 ? 2 : 3)
gcn_aggregate_union_kernel(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                           const float *__restrict__ val, const AggHubs hb, const AggTiles tl, unsigned hub_slots, int Nv,
                           const float *__restrict__ in, int B, int C, int L, const float *__restrict__ bias, int relu,
                           float *__restrict__ out, int BG, int n_tiles, int hubs_first, int ldi, int ldo, int nchunks,
                           int gc, int n_stages, int prefetch_next, int dense_mode) {
    // ldi / ldo: row strides (floats) of in / out; C channels are handled ([0, L) aggregated, [L, C) passed through)
    // nchunks x gc: the aggregated float4 groups are walked in nchunks chunks of gc (<= 32) groups
    extern __shared__ __align__(128) uint8_t au_dyn[];
    __shared__ __align__(16) uint32_t s_off[AG_WARPS][AT_STRIP];
    __shared__ __align__(16) float s_w[AG_WARPS][AT_STRIP];
    __shared__ __align__(16) float s_part[AG_WARPS][NG * 32 * 4];
    __shared__ uint32_t s_uoff[AU_MAXU];
    __shared__ __align__(8) uint64_t bars[8];
    pdl_launch_dependents();
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ngroups = C >> 2;
    const int gath = (L + 3) >> 2;
    const uint32_t row_bytes = (uint32_t)ldi * 4u;
    const size_t bstride = (size_t)Nv * ldi, bstride_o = (size_t)Nv * ldo;  // floats per batch element
    const int npass = ngroups - gath;       // pure pass-through groups
    const bool pass0 = lane < npass, pass1 = 32 + lane < npass;
    const int pv0 = (gath + lane) * 4, pv1 = (gath + 32 + lane) * 4;  // float offsets inside the row

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

    if (local < hub_slots) {
        // ------------------------------------------------------------------ hub CTAs (as in gcn_aggregate_tile_kernel)
        bool on[NG];
        uint32_t voff[NG];
#pragma unroll
        for (int n = 0; n < NG; ++n) {
            on[n] = lane + 32 * n < gath;
            voff[n] = (uint32_t)min(lane + 32 * n, gath - 1) * 16u;
        }
        auto do_row = [&](int i, const float *inb, float *outb, int beg, int end, int n4, float (&acc)[NG][4]) {
            const float *self = inb + (size_t)i * ldi;
            float *o = outb + (size_t)i * ldo;
            float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f), p1 = p0;
            if (pass0) p0 = __ldcs(reinterpret_cast<const float4 *>(self + pv0));
            if (pass1) p1 = __ldcs(reinterpret_cast<const float4 *>(self + pv1));
            const char *base[NG];
#pragma unroll
            for (int n = 0; n < NG; ++n) base[n] = reinterpret_cast<const char *>(inb) + voff[n];
            if (end - beg <= AT_STRIP) {
                strip_gather<NG>(s_off[warp], s_w[warp], n4, base, acc);
            } else {
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
        float acc[NG][4];
#pragma unroll
        for (int n = 0; n < NG; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
        const bool common = hb.n_common > 0;
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
            do_row(i, inb, outb, rb, re, n4, r);
        }
        return;
    }

    // ---------------------------------------------------------------------- tile CTAs: gather from the staged union
    const int tile = (int)(local - hub_slots);
    const int i0 = tile * AU_TV;
    const int b0 = (int)group * BG;
    const int nb = min(B, b0 + BG) - b0;
    const int u0 = tl.uptr[tile], U = tl.uptr[tile + 1] - u0;
    if (dense_mode) {
        // ------------------------------------------------------------------ dense-tile form (BG == 8: warp = batch element)
        // The tile's 8 rows as one small dense product  out[8 x C'] = A[8 x U] . X[U x C']  over the tile's neighbour
        // union: every union row is read from L2 ONCE per batch element (not once per edge) and feeds up to 8
        // accumulator rows held in registers; A (weights, 0 where a row does not use the column) sits in shared memory
        // as [U][8] and is read with two broadcast LDS.128 per union row.  A row's neighbours are still added in
        // ascending column order (the union is sorted), and fma(0, x, acc) == acc for finite x, so the result is
        // bit-identical to the sparse gather for finite inputs.
        float *sA = reinterpret_cast<float *>(au_dyn);
        __shared__ int s_valid[AG_WARPS];
        const int Up = (U + XW - 1) & ~(XW - 1);
        for (int e = threadIdx.x; e < Up * 8; e += AU_THREADS) sA[e] = 0.f;
        for (int u = threadIdx.x; u < Up; u += AU_THREADS) s_uoff[u] = u < U ? (uint32_t)tl.ucol[u0 + u] * row_bytes : 0u;
        __syncthreads();
        {
            const int i = i0 + warp;
            bool ok = false;
            if (i < Nv) {
                const int beg = rowptr[i], end = rowptr[i + 1];
                ok = !(hb.row_skip ? hb.row_skip[i] != 0 : (hub_slots > 0 && end - beg > HUB_DEG));
                if (ok)
                    for (int e = beg + lane; e < end; e += 32) sA[(int)tl.lidx[e] * 8 + warp] = val[e];
            }
            if (lane == 0) s_valid[warp] = ok ? 1 : 0;
        }
        __syncthreads();
        const int b = b0 + warp;
        if (b >= B) return;
        const float *inb = in + (size_t)b * bstride;
        float *outb = out + (size_t)b * bstride_o;
        if (npass > 0) {  // pass-through columns of the 8 rows: ask L2 for them now, copy them after the product
            const int line0 = (gath * 16) >> 7, lines = (ngroups * 16 + 127) >> 7;
            for (int r = 0; r < AU_TV; ++r)
                if (i0 + r < Nv && line0 + lane < lines)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(inb + (size_t)(i0 + r) * ldi) + (line0 + lane) * 128));
        }
        for (int ch = 0; ch < nchunks; ++ch) {
            const int g_here = min(gc, gath - ch * gc);
            const int gl = min(lane, g_here - 1);
            // accumulators as packed FP32x2 pairs (channels 0-1, 2-3): fma.rn.f32x2 is the same IEEE FMA per element as
            // fmaf -- the results stay bit-identical to the sparse gather -- but half the issue slots (ncu on the FFMA
            // version: issue slots 68 % busy, FMA pipe 50 %, L2 28 %: the product loop was issue-bound, not L2-bound)
            unsigned long long acc2[AU_TV][2];
#pragma unroll
            for (int r = 0; r < AU_TV; ++r) acc2[r][0] = acc2[r][1] = 0ull;
            const char *src = reinterpret_cast<const char *>(inb) + (size_t)(ch * gc + gl) * 16;
            for (int u = 0; u < Up; u += XW) {
                ulonglong2 x[XW];
#pragma unroll
                for (int j = 0; j < XW; ++j) x[j] = __ldg(reinterpret_cast<const ulonglong2 *>(src + s_uoff[u + j]));
#pragma unroll
                for (int j = 0; j < XW; ++j) {
                    const float4 w0 = *reinterpret_cast<const float4 *>(sA + (u + j) * 8);
                    const float4 w1 = *reinterpret_cast<const float4 *>(sA + (u + j) * 8 + 4);
                    const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                    for (int r = 0; r < AU_TV; ++r) {
                        unsigned long long w2;
                        asm("mov.b64 %0, {%1, %1};" : "=l"(w2) : "f"(w[r]));
                        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2[r][0]) : "l"(w2), "l"(x[j].x));
                        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2[r][1]) : "l"(w2), "l"(x[j].y));
                    }
                }
            }
            float acc[AU_TV][4];
#pragma unroll
            for (int r = 0; r < AU_TV; ++r) {
                acc[r][0] = __uint_as_float((unsigned)acc2[r][0]); acc[r][1] = __uint_as_float((unsigned)(acc2[r][0] >> 32));
                acc[r][2] = __uint_as_float((unsigned)acc2[r][1]); acc[r][3] = __uint_as_float((unsigned)(acc2[r][1] >> 32));
            }
            if (lane < g_here) {
                const int c0 = (ch * gc + lane) * 4;
                float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (bias) bv = *reinterpret_cast<const float4 *>(bias + c0);
#pragma unroll
                for (int r = 0; r < AU_TV; ++r) {
                    if (!s_valid[r]) continue;
                    const float *self = inb + (size_t)(i0 + r) * ldi;
                    float q[4] = {acc[r][0] + bv.x, acc[r][1] + bv.y, acc[r][2] + bv.z, acc[r][3] + bv.w};
                    if (!bias) { q[0] = acc[r][0]; q[1] = acc[r][1]; q[2] = acc[r][2]; q[3] = acc[r][3]; }
                    if (c0 + 4 > L) {
                        const float4 sv4 = *reinterpret_cast<const float4 *>(self + c0);
                        const float sv[4] = {sv4.x, sv4.y, sv4.z, sv4.w};
#pragma unroll
                        for (int t = 0; t < 4; ++t)
                            if (c0 + t >= L) q[t] = sv[t];
                    }
                    if (relu) {
#pragma unroll
                        for (int t = 0; t < 4; ++t) q[t] = fmaxf(q[t], 0.f);
                    }
                    __stcs(reinterpret_cast<float4 *>(outb + (size_t)(i0 + r) * ldo + c0), make_float4(q[0], q[1], q[2], q[3]));
                }
            }
        }
        if (npass > 0) {
            for (int r0 = 0; r0 < AU_TV; r0 += 2) {  // two rows (up to 4 float4 per lane) in flight
                float4 p[2][2];
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int r = r0 + rr;
                    const float *self = inb + (size_t)(i0 + r) * ldi;
                    const bool ok = s_valid[r] != 0;
                    p[rr][0] = (ok && pass0) ? __ldcs(reinterpret_cast<const float4 *>(self + pv0)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    p[rr][1] = (ok && pass1) ? __ldcs(reinterpret_cast<const float4 *>(self + pv1)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int r = r0 + rr;
                    if (!s_valid[r]) continue;
                    const float *self = inb + (size_t)(i0 + r) * ldi;
                    float *o = outb + (size_t)(i0 + r) * ldo;
                    if (pass0) __stcs(reinterpret_cast<float4 *>(o + pv0), relu4(p[rr][0], relu));
                    if (pass1) __stcs(reinterpret_cast<float4 *>(o + pv1), relu4(p[rr][1], relu));
                    for (int v = gath + 64 + lane; v < ngroups; v += 32)
                        __stcs(reinterpret_cast<float4 *>(o + v * 4), relu4(__ldcs(reinterpret_cast<const float4 *>(self + v * 4)), relu));
                }
            }
        }
        return;
    }
    const uint32_t srow = (uint32_t)gc * 16u;              // bytes of one union row in a stage
    const uint32_t stage_bytes = (uint32_t)U * srow;       // (the host sized the ring for the largest union)
    const uint32_t stage_pitch = (stage_bytes + 127u) & ~127u;
    const uint32_t ring = (au_smem_u32(au_dyn) + 127u) & ~127u;
    const uint32_t bar_full = au_smem_u32(&bars[0]), bar_empty = au_smem_u32(&bars[4]);
    if (threadIdx.x == 0) {
        for (int s = 0; s < n_stages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_full + 8 * s), "r"(AU_THREADS));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_empty + 8 * s), "r"(AG_WARPS));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int u = threadIdx.x; u < U; u += AU_THREADS) s_uoff[u] = (uint32_t)tl.ucol[u0 + u] * row_bytes;
    // this warp's row: (local byte offset in a stage, weight) list in the warp's strip
    const int i = i0 + warp;
    int n4 = -1;  // -1: no row here / a hub CTA's row
    if (i < Nv) {
        const int beg = rowptr[i], end = rowptr[i + 1];
        const bool skip = hb.row_skip ? hb.row_skip[i] != 0 : (hub_slots > 0 && end - beg > HUB_DEG);
        if (!skip) {
            const int cnt = end - beg;  // <= AT_STRIP: the host checked every non-hub row
            n4 = (cnt + 3) & ~3;
            for (int e = lane; e < n4; e += 32) {
                const bool real = e < cnt;
                s_off[warp][e] = real ? (uint32_t)tl.lidx[beg + e] * srow : 0u;
                s_w[warp][e] = real ? val[beg + e] : 0.f;
            }
        }
    }
    __syncthreads();

    const int total = nb * nchunks;  // iteration k: batch element k / nchunks, column chunk k % nchunks
    // Loads of iteration j: warp w copies the union rows u = w, w + 8, ... (one row per instruction, lanes < groups copy
    // 16 bytes each) into slot j % n_stages and reports completion to the slot's mbarrier.
    auto load_iter = [&](int j) {
        const int s = j % n_stages;
        // the slot's previous use (iteration j - n_stages) must have been read by every warp
        if (j >= n_stages) au_mbar_wait(bar_empty + 8 * s, ((j / n_stages) - 1) & 1);
        const int bb = j / nchunks, ch = j - bb * nchunks;
        const int g_here = min(gc, gath - ch * gc);
        const char *src = reinterpret_cast<const char *>(in + (size_t)(b0 + bb) * bstride) + (size_t)ch * gc * 16 + lane * 16;
        const uint32_t dst = ring + (uint32_t)s * stage_pitch + (uint32_t)lane * 16u;
        if (lane < g_here) {
#pragma unroll 2
            for (int u = warp; u < U; u += AG_WARPS)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)u * srow), "l"(src + s_uoff[u]) : "memory");
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar_full + 8 * s) : "memory");
    };
    const int ahead = n_stages - 2 > 0 ? n_stages - 2 : 1;  // one slot of slack: a warp may run one iteration ahead of the slowest
    for (int j = 0; j < ahead && j < total; ++j) load_iter(j);
    for (int k = 0; k < total; ++k) {
        const int s = k % n_stages;
        const int bb = k / nchunks, ch = k - bb * nchunks;
        if (k + ahead < total) load_iter(k + ahead);
        const float *inb = in + (size_t)(b0 + bb) * bstride;
        float *outb = out + (size_t)(b0 + bb) * bstride_o;
        const float *self = inb + (size_t)i * ldi;
        float *o = outb + (size_t)i * ldo;
        float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f), p1 = p0;
        const bool row_here = i < Nv;
        if (ch == 0 && row_here) {
            // pass-through part of the row (requested before the wait); the next batch element's copy of it is asked
            // from L2 one element ahead (its head columns arrive with the union copies anyway)
            if (pass0) p0 = __ldcs(reinterpret_cast<const float4 *>(self + pv0));
            if (pass1) p1 = __ldcs(reinterpret_cast<const float4 *>(self + pv1));
            if (prefetch_next && bb + 1 < nb) {
                const int line0 = (gath * 16) >> 7, lines = (ngroups * 16 + 127) >> 7;
                if (line0 + lane < lines)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(self + bstride) + (line0 + lane) * 128));
            }
        }
        au_mbar_wait(bar_full + 8 * s, (k / n_stages) & 1);
        if (n4 >= 0) {
            const int g_here = min(gc, gath - ch * gc);
            const int gl = min(lane, g_here - 1);  // lanes beyond the chunk are clamped onto its last group
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            strip_gather_smem(s_off[warp], s_w[warp], n4, ring + (uint32_t)s * stage_pitch + (uint32_t)gl * 16u, acc);
            if (lane < g_here) {
                const int c0 = (ch * gc + lane) * 4;
                float r[4] = {acc[0], acc[1], acc[2], acc[3]};
                if (bias) {
                    const float4 bv = *reinterpret_cast<const float4 *>(bias + c0);
                    r[0] += bv.x; r[1] += bv.y; r[2] += bv.z; r[3] += bv.w;
                }
                if (c0 + 4 > L) {  // boundary group: channels >= L pass through
                    const float4 sv4 = *reinterpret_cast<const float4 *>(self + c0);
                    const float sv[4] = {sv4.x, sv4.y, sv4.z, sv4.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (c0 + q >= L) r[q] = sv[q];
                }
                if (relu) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) r[q] = fmaxf(r[q], 0.f);
                }
                __stcs(reinterpret_cast<float4 *>(o + c0), make_float4(r[0], r[1], r[2], r[3]));
            }
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_empty + 8 * s) : "memory");
        if (ch == 0 && row_here) {
            // (hub rows' pass-through columns are written by their hub CTA)
            if (n4 >= 0) {
                if (pass0) __stcs(reinterpret_cast<float4 *>(o + pv0), relu4(p0, relu));
                if (pass1) __stcs(reinterpret_cast<float4 *>(o + pv1), relu4(p1, relu));
                for (int v = gath + 64 + lane; v < ngroups; v += 32)
                    __stcs(reinterpret_cast<float4 *>(o + v * 4), relu4(__ldcs(reinterpret_cast<const float4 *>(self + v * 4)), relu));
            }
        }
    }
}

static unsigned long long g_au_optin[6] = {0ull, 0ull, 0ull, 0ull, 0ull, 0ull};

// Returns 1 when the union kernel was launched, 0 when the shape is outside its range (the caller falls back to the
// tile kernel), < 0 on error.
int aggregate_union_launch(int mode, const int32_t *rowptr, const int32_t *col, const float *val, const AggHubs &hb,
                           const AggTiles &tl, int max_union, bool have_hubs, int64_t Nv, const float *in, int64_t B, int64_t C,
                           int64_t L, const float *bias, int relu, float *out, int64_t ldi, int64_t ldo, int hubs_first,
                           int prefetch_next, cudaStream_t st) {
    const int dense_mode = mode == 1 ? 1 : 0;  // 1: dense-tile product (warp = batch element), else the shared-memory ring
    const int gath = (int)((L + 3) / 4);
    if (max_union <= 0 || max_union > AU_MAXU || gath < 1 || gath > 96) return 0;
    const int nchunks = (gath + 31) / 32;
    const int gc = (gath + nchunks - 1) / nchunks;
    const size_t pitch = (((size_t)max_union * gc * 16) + 127) & ~(size_t)127;
    // ring depth: as many slots (<= 4) as still let three CTAs share an SM (227 KB minus ~13-21 KB static each), >= 2
    const int NG = gath <= 32 ? 1 : (gath <= 64 ? 2 : 3);
    const size_t stat = 9 * 1024 + (size_t)NG * 4096 + 1200;
    int n_stages = 4;
    while (n_stages > 2 && 3 * (stat + n_stages * pitch + 128) > 225 * 1024) --n_stages;
    if (PTK_TUNING_ENV("PTK_AGG_STAGES") > 0) n_stages = PTK_TUNING_ENV("PTK_AGG_STAGES");
    size_t smem = (size_t)n_stages * pitch + 128;
    if (dense_mode) smem = (size_t)((max_union + 7) & ~7) * 8 * sizeof(float) + 128;  // A [U][8]
    if (stat + smem > 220 * 1024) return 0;
    int BG = 8;  // the dense-tile form needs exactly 8: warp = batch element
    const long long slots = 6LL * sm_count();
    const int n_tiles = (int)ceil_div(Nv, AU_TV);
    if (!dense_mode) {
        while (BG > 1 && (long long)n_tiles * ceil_div(B, BG) < slots) BG >>= 1;
        if (PTK_TUNING_ENV("PTK_AGG_BG") > 0) BG = PTK_TUNING_ENV("PTK_AGG_BG");
    }
    const bool common = hb.n_common > 0;
    const unsigned hub_slots = have_hubs ? (unsigned)(common ? BG : BG * hb.n_hubs) : 0u;
    const unsigned grid = (unsigned)((hub_slots + n_tiles) * ceil_div(B, BG));
    int dev = 0;
    PTK_CHECK_CUDA(cudaGetDevice(&dev));
    const bool xw8 = dense_mode && B <= 32;
#define PTK_UNION(NGv, XWv)                                                                                               \
    do {                                                                                                                  \
        const int slot = (NGv - 1) * 2 + (XWv == 8 ? 1 : 0);                                                              \
        if (dev >= 64 || !((g_au_optin[slot] >> dev) & 1ull)) {                                                           \
            PTK_CHECK_CUDA(cudaFuncSetAttribute(gcn_aggregate_union_kernel<NGv, XWv>,                                      \
                                                cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));                \
            if (dev < 64) g_au_optin[slot] |= 1ull << dev;                                                                \
        }                                                                                                                 \
        launch_pdl(gcn_aggregate_union_kernel<NGv, XWv>, dim3(grid), dim3(AU_THREADS), smem, st, rowptr, col, val, hb, tl,  \
                   hub_slots, (int)Nv, in, (int)B, (int)C, (int)L, bias, relu, out, BG, n_tiles, hubs_first, (int)ldi,   \
                   (int)ldo, nchunks, gc, n_stages, prefetch_next, dense_mode);                                           \
    } while (0)
    if (NG == 1) { if (xw8) PTK_UNION(1, 8); else PTK_UNION(1, 4); }
    else if (NG == 2) { if (xw8) PTK_UNION(2, 8); else PTK_UNION(2, 4); }
    else { if (xw8) PTK_UNION(3, 8); else PTK_UNION(3, 4); }
#undef PTK_UNION
    PTK_CHECK_LAUNCH();
    return 1;
}

}  // namespace ptk
