// GCN per-vertex linear layers on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// Replaces `torch.matmul(features, self.weight)` (pterotactyl/reconstruction/vision/model.py:352) and
// its dgrad where the layer is a true GEMM (300x300, 448x300; K >= 32).  The parity contract is FP32
// (1e-5 relative), so a single TF32 pass (10-bit mantissa) is not admissible; this kernel runs the
// error-compensated 3xTF32 scheme:
//       a = a_hi + a_lo,  a_hi = tf32_rna(a),  a_lo = a - a_hi (truncated to TF32 by the tensor core)   (same for b)
//       D = a_hi*b_lo + a_lo*b_hi + a_hi*b_hi        (a_lo*b_lo ~ 2^-22 relative is dropped)
// CHUNKED ACCUMULATION.  The tensor core adds into its FP32 accumulator with truncation: measured here,
// one accumulator carried over the whole reduction gives an error that grows LINEARLY with K (2.4e-6 at
// K=300, 4x a scalar FP32 loop) and, being biased, linearly with network depth (3.6e-5 on the input
// gradient of the 20-layer GCN).  So each TMEM accumulator only integrates ONE 32-wide k-block (the 8
// small correction products first, then the 4 hi*hi products); a dedicated group of warps drains it
// (tcgen05.ld) and adds it into FP32 registers with round-to-nearest while the tensor core fills the
// next TMEM buffer.  Result: error at the level of a scalar FP32 loop, unbiased.
//
//   D (M x N) = A (M x K, row-major: K contiguous)  .  B^T   with B stored (N x K), K contiguous
//     fwd   : A = X,  B = W^T (pre-transposed + pre-split, tiny)          D = H
//     dgrad : A = gH, B = W   (as stored: (K_in x N_out) is "N x K")      D = gX  [* (act > 0)]
//
// One CTA = one 128 x 160 output tile, 512 threads, 3-stage smem ring, 3 TMEM accumulator buffers:
//   warp 0       TMA producer: raw FP32 A tile (128 x 32) + pre-split B_hi/B_lo tiles (160 x 32) per
//                k-block, SWIZZLE_128B, mbarrier expect_tx.  (Loading B raw and splitting it in the kernel
//                too -- 36 instead of 56 KB per k-block -- was measured: not faster, the split of B then
//                sits on the critical path of every k-block.)
//   warp 1       MMA issuer (one elected lane): 12 x tcgen05.mma.kind::tf32 (M128 N160 K8) per k-block;
//                tcgen05.commit releases the smem stage and publishes the TMEM buffer
//   warps 4-7    converter: split the raw A tile in place into a_hi (overwrites raw) and a_lo
//                (elementwise, position preserving => swizzle-agnostic), fence.proxy.async, arrive
//   warps 8-15   drain + epilogue: per k-block tcgen05.ld the finished buffer, acc += chunk (RN);
//                at the end optional ReLU mask and 128-bit stores (thread = one row x 80 columns)
#include <cuda.h>
#include <stdlib.h>

#include "ptk_common.cuh"

namespace ptk {

constexpr int TG_BM = 128;        // rows per CTA (UMMA M, cta_group::1)
constexpr int TG_BN = 160;        // columns per CTA (UMMA N)
constexpr int TG_BK = 32;         // fp32 elements per k-block = 128 B = one SWIZZLE_128B atom row
constexpr int TG_STAGES = 3;      // smem ring: 3 x 72 KB
constexpr int TG_NBUF = 3;        // TMEM accumulator buffers: 3 x 160 columns
constexpr int TG_THREADS = 512;
constexpr int TG_CONV_WARPS = 4;  // warps 4-7
constexpr int TG_A_BYTES = TG_BM * TG_BK * 4;   // 16 KB
constexpr int TG_B_BYTES = TG_BN * TG_BK * 4;   // 20 KB
constexpr int TG_STAGE_BYTES = 2 * TG_A_BYTES + 2 * TG_B_BYTES;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// the same load delivered to the CTAs of `mask` (same shared-memory offset and mbarrier offset in each of them)
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1,
                                               uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// completion of the MMAs issued so far reported to the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}
// hi = x rounded to TF32 (nearest, ties away -- what cvt.rna.tf32.f32 returns for finite x) in two integer
// instructions; the PTX cvt expands to ~4 ALU instructions with its NaN/Inf handling and made the converter
// warps the slowest stage of the pipeline.  A NaN either stays a NaN or becomes -0 here; its lo part
// (x - hi) is a NaN in both cases, so non-finite inputs still give non-finite outputs.
__device__ __forceinline__ uint32_t tf32_hi(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
// lo = x - hi exactly (FP32); the tensor core reads its upper 19 bits, i.e. truncates it to TF32: an error of
// <= 2^-21 |x| with the sign of lo, i.e. random -- same order as the dropped lo*lo term.
__device__ __forceinline__ void split_tf32(float x, uint32_t &h, uint32_t &l) {
    h = tf32_hi(x);
    l = __float_as_uint(x - __uint_as_float(h));
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4 &v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 "version 1"):
//   start address >> 4 | LBO (unused for swizzled K-major) | SBO = 8 rows * 128 B = 1024 B | layout 2
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;                    // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset: next 8-row core-matrix group
    d |= (uint64_t)1 << 46;                    // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
    return d;
}

// kind::tf32 instruction descriptor: FP32 accumulate, A/B = TF32, both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t make_idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TG_BM >> 4) << 24);
}

constexpr int TG_SPLIT_BATCH = 32;  // layers per batched weight-split launch

struct TGParams {
    int M, N, K;          // D is M x N, reduction K
    float *D;
    const uint32_t *mask; // optional ReLU mask of D, packed: bit (c & 31) of word [row * wpr + (c >> 5)] = act[row, c] > 0
    int wpr;              // mask words per row = ceil(N / 32)
    int dbg;              // development switches (PTK_TG_DEBUG): 1 = skip A split, 2 = skip drain loads, 4 = skip MMAs
    float *part;          // (gridDim.x, 128 x 160) partial tiles of the CTAs whose range starts inside a tile
    int *flags;           // (gridDim.x) 1 once part[c] is complete (zeroed by the weight-split kernel before)
};

// PAIR: launched as clusters of two CTAs that work on the two 128-row tiles of a 256-row pair with the SAME column tile
// and k-block sequence.  Each CTA loads its own A tile and HALF of the split weight tiles (80 of the 160 rows of B_hi and
// B_lo), multicast to both: 36 instead of 56 KB per CTA and k-block come out of L2.  A stage is reused once BOTH CTAs'
// MMAs have read it (tcgen05.commit multicast onto both empty barriers).
template <bool MASK, bool PAIR>
__global__ void __launch_bounds__(TG_THREADS, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_bhi,
                   const __grid_constant__ CUtensorMap map_blo, const TGParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // per stage: [A_hi (raw) 16 KB][A_lo 16 KB][B_hi 20 KB][B_lo 20 KB], every piece 1024-B aligned
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bars[3 * TG_STAGES + 2 * TG_NBUF];
    __shared__ uint32_t tmem_base_slot;
    const uint32_t bar_full = smem_u32(&bars[0]);                  // TMA landed         (1 + tx)
    const uint32_t bar_conv = smem_u32(&bars[TG_STAGES]);          // A split finished   (4 warps)
    const uint32_t bar_empty = smem_u32(&bars[2 * TG_STAGES]);     // MMAs done with smem (commit)
    const uint32_t bar_tfull = smem_u32(&bars[3 * TG_STAGES]);     // TMEM buffer complete (commit)
    const uint32_t bar_tempty = smem_u32(&bars[3 * TG_STAGES + TG_NBUF]);  // TMEM buffer drained (8 warps)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = (p.K + TG_BK - 1) / TG_BK;
    pdl_launch_dependents();  // the successor may start its own prologue while this grid runs (ptk_common.cuh)
    // persistent, work split by k-blocks ("stream-K"): the (tile, k-block) pairs are numbered tile-major and CTA c
    // takes the contiguous range [c U / G, (c+1) U / G) of them -- 488 tiles on 148 SMs would otherwise cost 4 tile
    // times for 3.3 tiles of work per SM.  A range is at least one tile long, so a tile is shared by at most two
    // CTAs.  Every CTA walks its range from the END: CTA c therefore starts with the tile it shares with CTA c+1
    // (the tile's first k-blocks), stores that partial tile to its workspace slot and raises its flag; CTA c+1 gets
    // to the tile's remaining k-blocks LAST, adds the partial (head + tail: fixed order => deterministic) and runs
    // the epilogue.  A CTA thus only ever waits for a lower-numbered CTA, which was dispatched before it.
    // tile t -> (m-tile t / tiles_n, n-tile t % tiles_n).  All pipeline counters (smem stage, TMEM buffer) run
    // across tiles, so the epilogue stores of one tile overlap the TMA / MMA work of the next.
    const int tiles_n = (p.N + TG_BN - 1) / TG_BN;
    const int tiles_m = (p.M + TG_BM - 1) / TG_BM;
    // PAIR: the work items are (256-row pair, column tile) x k-block, dealt out to the CLUSTERS; rank r of a cluster takes
    // the pair's r-th 128-row tile (rows beyond M: TMA zero-fill, no stores)
    const int crank = PAIR ? (int)cluster_ctarank() : 0;
    const int step = PAIR ? 2 : 1;                                   // predecessor / successor CTA distance
    const int workers = PAIR ? (int)gridDim.x / 2 : (int)gridDim.x;
    const int me = PAIR ? (int)blockIdx.x / 2 : (int)blockIdx.x;
    const int num_tiles = (PAIR ? (tiles_m + 1) / 2 : tiles_m) * tiles_n;
    const long long units = (long long)num_tiles * num_kb;
    const int u0 = (int)(units * me / workers), u1 = (int)(units * (me + 1) / workers);
    auto tile_m0 = [&](int tile) { return ((PAIR ? 2 * (tile / tiles_n) + crank : tile / tiles_n)) * TG_BM; };

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_bhi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_blo) : "memory");
        for (int s = 0; s < TG_STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_conv + 8 * s, TG_CONV_WARPS);
            mbar_init(bar_empty + 8 * s, PAIR ? 2 : 1);
        }
        for (int b = 0; b < TG_NBUF; ++b) {
            mbar_init(bar_tfull + 8 * b, 1);
            mbar_init(bar_tempty + 8 * b, 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: all 512 columns (1 CTA per SM by shared-memory size)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;
    if (PAIR) cluster_sync_all();  // the peer's barriers exist before anything is multicast onto them
    pdl_wait();  // barriers + TMEM are set up; from here on global memory of the predecessors is read

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int it = 0;
            for (int hi = u1; hi > u0;) {
                const int tile = (hi - 1) / num_kb;
                const int kb_begin = max(u0 - tile * num_kb, 0), kb_end = hi - tile * num_kb;
                hi = tile * num_kb + kb_begin;
                const int m0 = tile_m0(tile), n_base = (tile % tiles_n) * TG_BN;
                for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
                    const int s = it % TG_STAGES;
                    const uint32_t ph = (it / TG_STAGES) & 1;
                    mbar_wait(bar_empty + 8 * s, ph ^ 1);
                    const uint32_t sa = smem_u32(smem + (size_t)s * TG_STAGE_BYTES);
                    mbar_expect_tx(bar_full + 8 * s, TG_A_BYTES + 2 * TG_B_BYTES);
                    tma_load_2d(sa, &map_a, bar_full + 8 * s, kb * TG_BK, m0);
                    if (PAIR) {
                        // this CTA's half of the weight tiles (80 rows = 10 swizzle atoms of 1 KB), delivered to both CTAs
                        const uint32_t hoff = (uint32_t)crank * (TG_B_BYTES / 2);
                        tma_load_2d_mc(sa + 2 * TG_A_BYTES + hoff, &map_bhi, bar_full + 8 * s, kb * TG_BK,
                                       n_base + crank * (TG_BN / 2), 3);
                        tma_load_2d_mc(sa + 2 * TG_A_BYTES + TG_B_BYTES + hoff, &map_blo, bar_full + 8 * s, kb * TG_BK,
                                       n_base + crank * (TG_BN / 2), 3);
                    } else {
                        tma_load_2d(sa + 2 * TG_A_BYTES, &map_bhi, bar_full + 8 * s, kb * TG_BK, n_base);
                        tma_load_2d(sa + 2 * TG_A_BYTES + TG_B_BYTES, &map_blo, bar_full + 8 * s, kb * TG_BK, n_base);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = make_idesc_tf32(TG_BN);
        const int total_kb = u1 - u0;
        for (int kb = 0; kb < total_kb; ++kb) {
            const int s = kb % TG_STAGES, b = kb % TG_NBUF;
            mbar_wait(bar_full + 8 * s, (kb / TG_STAGES) & 1);        // B tiles landed
            mbar_wait(bar_conv + 8 * s, (kb / TG_STAGES) & 1);        // A split done
            mbar_wait(bar_tempty + 8 * b, ((kb / TG_NBUF) & 1) ^ 1);  // accumulator buffer drained
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0 && !(p.dbg & 4)) {
                const uint32_t sa = smem_u32(smem + (size_t)s * TG_STAGE_BYTES);
                const uint64_t d_ahi = make_kmajor_sw128_desc(sa);
                const uint64_t d_alo = make_kmajor_sw128_desc(sa + TG_A_BYTES);
                const uint64_t d_bhi = make_kmajor_sw128_desc(sa + 2 * TG_A_BYTES);
                const uint64_t d_blo = make_kmajor_sw128_desc(sa + 2 * TG_A_BYTES + TG_B_BYTES);
                const uint32_t d_tmem = tmem_base + (uint32_t)(b * TG_BN);
                // small correction products first (into a cleared accumulator), then the hi*hi products
#pragma unroll
                for (int ks = 0; ks < TG_BK / 8; ++ks) {
                    const uint64_t ko = (uint64_t)(ks * 32 >> 4);  // +32 B per k-step inside the swizzle atom
                    umma_tf32(d_tmem, d_ahi + ko, d_blo + ko, idesc, ks != 0);
                    umma_tf32(d_tmem, d_alo + ko, d_bhi + ko, idesc, 1);
                }
#pragma unroll
                for (int ks = 0; ks < TG_BK / 8; ++ks) {
                    const uint64_t ko = (uint64_t)(ks * 32 >> 4);
                    umma_tf32(d_tmem, d_ahi + ko, d_bhi + ko, idesc, 1);
                }
                // smem stage free once these MMAs have read it (PAIR: the peer may then overwrite its half too)
                if (PAIR) umma_commit_mc(bar_empty + 8 * s, 3); else umma_commit(bar_empty + 8 * s);
                umma_commit(bar_tfull + 8 * b);   // accumulator chunk complete
            } else if (lane == 0) {
                if (PAIR) umma_commit_mc(bar_empty + 8 * s, 3); else umma_commit(bar_empty + 8 * s);
                umma_commit(bar_tfull + 8 * b);
            }
            __syncwarp();
        }
    } else if (warp >= 4 && warp < 4 + TG_CONV_WARPS) {
        // ===================== converter warps =====================
        const int ct = threadIdx.x - 128;  // 0..127
        const int total_kb = u1 - u0;
        for (int kb = 0; kb < total_kb; ++kb) {
            const int s = kb % TG_STAGES;
            mbar_wait(bar_full + 8 * s, (kb / TG_STAGES) & 1);
            const uint32_t stage = smem_u32(smem + (size_t)s * TG_STAGE_BYTES);
            constexpr int CA = TG_A_BYTES / 16;
            static_assert(CA % (32 * TG_CONV_WARPS) == 0, "converter chunks must divide evenly");
#pragma unroll
            for (int i = 0; i < CA / (32 * TG_CONV_WARPS); ++i) {
                if (p.dbg & 1) break;
                const uint32_t hi = stage + 16u * (uint32_t)(ct + i * 32 * TG_CONV_WARPS);  // raw -> a_hi in place
                const uint32_t lo = hi + (uint32_t)TG_A_BYTES;
                const uint4 v = lds128(hi);
                uint4 h, l;
                split_tf32(__uint_as_float(v.x), h.x, l.x);
                split_tf32(__uint_as_float(v.y), h.y, l.y);
                split_tf32(__uint_as_float(v.z), h.z, l.z);
                split_tf32(__uint_as_float(v.w), h.w, l.w);
                sts128(hi, h);
                sts128(lo, l);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> async proxy (UMMA)
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_conv + 8 * s);
        }
    } else if (warp >= 8) {
        // ===================== drain + epilogue warps =====================
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int half = (warp - 8) >> 2;       // column half: [0,80) or [80,160)
        int it = 0;                             // running chunk index across tiles
        const int dt = threadIdx.x - 256;       // 0..255 among the drain threads
        for (int hi = u1; hi > u0;) {
        const int tile = (hi - 1) / num_kb;
        const int kb_begin = max(u0 - tile * num_kb, 0), kb_end = hi - tile * num_kb;
        hi = tile * num_kb + kb_begin;
        const bool head_part = kb_end < num_kb;    // the tile's last k-blocks belong to CTA blockIdx.x + 1: publish
        const bool tail_part = kb_begin > 0;       // the tile's first k-blocks came from CTA blockIdx.x - 1: combine
        const int m0 = tile_m0(tile), n_base = (tile % tiles_n) * TG_BN;
        const int row = m0 + q * 32 + lane;
        float acc[80];
#pragma unroll
        for (int j = 0; j < 80; ++j) acc[j] = 0.f;
        // ReLU mask of this thread's 80 outputs, fetched while the pipeline fills (bit j: act > 0)
        uint32_t mbits[3] = {0u, 0u, 0u};
        if (MASK && row < p.M && !head_part) {
            // 80 mask bits of this thread's outputs from the packed row mask (4 word loads instead of 20 strided
            // 16-byte loads of the activation itself: the epilogue is one thread per row)
            const int nb0 = n_base + half * 80;
            const int w0 = nb0 >> 5, sh = nb0 & 31;
            const uint32_t *mr = p.mask + (size_t)row * p.wpr;
            uint32_t w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) w[i] = (w0 + i < p.wpr) ? __ldg(mr + w0 + i) : 0u;
#pragma unroll
            for (int i = 0; i < 3; ++i) mbits[i] = __funnelshift_r(w[i], w[i + 1], sh);
        }
        for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
            const int b = it % TG_NBUF;
            mbar_wait(bar_tfull + 8 * b, (it / TG_NBUF) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * TG_BN + half * 80);
#pragma unroll
            for (int c0 = 0; c0 < 80; c0 += 16) {
                if (p.dbg & 2) break;
                uint32_t r[16];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(taddr + (uint32_t)c0));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[c0 + j] += __uint_as_float(r[j]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * b);
        }
        if (head_part) {
            // partial tile -> slot of this CTA as [20 float4 per thread][256 threads] (coalesced), then the flag
            float4 *slot = reinterpret_cast<float4 *>(p.part) + (size_t)blockIdx.x * (TG_BM * TG_BN / 4);
#pragma unroll
            for (int v = 0; v < 20; ++v)
                __stcg(slot + v * 256 + dt, make_float4(acc[4 * v], acc[4 * v + 1], acc[4 * v + 2], acc[4 * v + 3]));
            __threadfence();
            asm volatile("bar.sync 1, 256;" ::: "memory");  // the 8 drain warps
            if (dt == 0) {
                __threadfence();
                *reinterpret_cast<volatile int *>(p.flags + blockIdx.x) = 1;
            }
            continue;
        }
        if (tail_part) {
            const volatile int *flag = p.flags + blockIdx.x - step;
            while (*flag == 0) __nanosleep(64);
            __threadfence();
            const float4 *slot = reinterpret_cast<const float4 *>(p.part) + (size_t)(blockIdx.x - step) * (TG_BM * TG_BN / 4);
#pragma unroll
            for (int v = 0; v < 20; ++v) {
                const float4 t = __ldcg(slot + v * 256 + dt);
                acc[4 * v] = t.x + acc[4 * v]; acc[4 * v + 1] = t.y + acc[4 * v + 1];
                acc[4 * v + 2] = t.z + acc[4 * v + 2]; acc[4 * v + 3] = t.w + acc[4 * v + 3];
            }
        }
        // epilogue: this thread owns row `row`, columns n_base + half*80 + [0, 80)
        const int nb = n_base + half * 80;
        if (row < p.M) {
            float *dst = p.D + (size_t)row * p.N + nb;
            if (MASK) {
#pragma unroll
                for (int j = 0; j < 80; ++j) acc[j] = (mbits[j >> 5] >> (j & 31)) & 1u ? acc[j] : 0.f;
            }
            if ((p.N & 3) == 0) {
#pragma unroll
                for (int v = 0; v < 20; ++v)
                    if (nb + 4 * v + 4 <= p.N)
                        *reinterpret_cast<float4 *>(dst + 4 * v) =
                            make_float4(acc[4 * v], acc[4 * v + 1], acc[4 * v + 2], acc[4 * v + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < 80; ++j)
                    if (nb + j < p.N) dst[j] = acc[j];
            }
        }
        }  // tile loop
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (PAIR) cluster_sync_all();  // no CTA leaves while its peer may still signal its barriers
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// =================================================================================================
// 2-SM form of the kernel above: cta_group::2 MMAs.
//
// What the probes showed (tools/tc_gemm_probe.py): the single-CTA kernel sits on its operand floor (273 MB of tiles
// through L2 -> shared memory), and the multicast pair form removes a third of those bytes but gains nothing because
// three 72 KB stages -- all of shared memory -- do not cover the TMA -> split -> MMA -> release chain.  Here the two CTAs
// of a cluster form ONE 256 x 160 tile: each keeps its 128 rows of A (raw -> a_hi in place, a_lo) and only HALF of the
// split weight tile (80 of the 160 rows of B_hi / B_lo); the leader CTA issues tcgen05.mma.cta_group::2 (M = 256), which
// reads A from both CTAs' shared memory and the two B halves from their owners, and writes each CTA's 128 accumulator
// rows into that CTA's own TMEM.  A stage is 52 KB: FOUR stages fit, and 36 instead of 56 KB per CTA and k-block come
// out of L2 AND into the SM.  Everything downstream (drain warps adding each k-block's chunk in FP32 registers, the
// k-block work split with partial tiles and flags, the masked epilogue) is per CTA exactly as above.
// MEASURED (PTK_TG_PAIR=3, M = 31184, K = N = 300): same results (2.2e-7 vs fp64), but 62.1 us against 53.0 us for the
// single-CTA kernel, and 52.2 us with split, MMAs and drain all disabled: every k-block now crosses the pair three times
// (the peer's B half and split report to the leader, the leader's commit releases the peer, both drains report back) and
// the accumulator ring is still 3 deep (3 x 160 of 512 TMEM columns), so the fourth shared-memory stage cannot fill.
// With 32-wide k-blocks that chain costs more than the saved bytes.  Kept selectable, off by default.
//   barriers (same offsets in both CTAs; "leader's" = the copy in CTA rank 0, reached through mapa):
//     bar_a[s]      local     own A tile landed (TMA)                          -> the CTA's converter warps
//     bar_b[s]      leader's  both B halves landed (each CTA's TMA signals it) -> leader's MMA warp
//     bar_conv[s]   leader's  A split done in both CTAs (2 x 4 warps)          -> leader's MMA warp
//     bar_empty[s]  local     MMAs have read the stage (commit, multicast)     -> the CTA's TMA producer
//     bar_tfull[b]  local     accumulator chunk complete (commit, multicast)   -> the CTA's drain warps
//     bar_tempty[b] leader's  chunk drained in both CTAs (2 x 8 warps)         -> leader's MMA warp
constexpr int T2_STAGES = 4;
constexpr int T2_BH_BYTES = TG_B_BYTES / 2;                           // 10 KB: 80 rows of B_hi (or B_lo)
constexpr int T2_STAGE_BYTES = 2 * TG_A_BYTES + 2 * T2_BH_BYTES;      // 52 KB

__device__ __forceinline__ uint32_t mapa_rank0(uint32_t addr) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(addr));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load into THIS CTA's shared memory whose bytes are counted on a barrier given by its cluster address (the leader's)
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap *map, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}
// kind::tf32 instruction descriptor of the pair: M = 256 (128 rows per CTA), N = n
__host__ __device__ constexpr uint32_t make_idesc_tf32_m256(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

template <bool MASK>
__global__ void __launch_bounds__(TG_THREADS, 1)
gemm_tf32x3_2sm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_bhi,
                       const __grid_constant__ CUtensorMap map_blo, const TGParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // per stage: [A_hi (raw) 16 KB][A_lo 16 KB][B_hi half 10 KB][B_lo half 10 KB], every piece 1024-B aligned
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bars[4 * T2_STAGES + 2 * TG_NBUF];
    __shared__ uint32_t tmem_base_slot;
    const uint32_t bar_a = smem_u32(&bars[0]);
    const uint32_t bar_b = smem_u32(&bars[T2_STAGES]);
    const uint32_t bar_conv = smem_u32(&bars[2 * T2_STAGES]);
    const uint32_t bar_empty = smem_u32(&bars[3 * T2_STAGES]);
    const uint32_t bar_tfull = smem_u32(&bars[4 * T2_STAGES]);
    const uint32_t bar_tempty = smem_u32(&bars[4 * T2_STAGES + TG_NBUF]);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = (p.K + TG_BK - 1) / TG_BK;
    pdl_launch_dependents();
    const int tiles_n = (p.N + TG_BN - 1) / TG_BN;
    const int tiles_m = (p.M + TG_BM - 1) / TG_BM;
    const int crank = (int)cluster_ctarank();
    const bool leader = crank == 0;
    const int workers = (int)gridDim.x / 2, me = (int)blockIdx.x / 2;
    const int num_tiles = ((tiles_m + 1) / 2) * tiles_n;            // (256-row pair, column tile)
    const long long units = (long long)num_tiles * num_kb;
    const int u0 = (int)(units * me / workers), u1 = (int)(units * (me + 1) / workers);
    auto tile_m0 = [&](int tile) { return (2 * (tile / tiles_n) + crank) * TG_BM; };

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_bhi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_blo) : "memory");
        for (int s = 0; s < T2_STAGES; ++s) {
            mbar_init(bar_a + 8 * s, 1);
            mbar_init(bar_b + 8 * s, 1);
            mbar_init(bar_conv + 8 * s, 2 * TG_CONV_WARPS);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int b = 0; b < TG_NBUF; ++b) {
            mbar_init(bar_tfull + 8 * b, 1);
            mbar_init(bar_tempty + 8 * b, 2 * 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: all 512 columns in both CTAs of the pair
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;
    cluster_sync_all();  // both CTAs' barriers and TMEM exist before anything is signalled across
    pdl_wait();

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            int it = 0;
            for (int hi = u1; hi > u0;) {
                const int tile = (hi - 1) / num_kb;
                const int kb_begin = max(u0 - tile * num_kb, 0), kb_end = hi - tile * num_kb;
                hi = tile * num_kb + kb_begin;
                const int m0 = tile_m0(tile), n_half = (tile % tiles_n) * TG_BN + crank * (TG_BN / 2);
                for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
                    const int s = it % T2_STAGES;
                    const uint32_t ph = (it / T2_STAGES) & 1;
                    mbar_wait(bar_empty + 8 * s, ph ^ 1);
                    const uint32_t sa = smem_u32(smem + (size_t)s * T2_STAGE_BYTES);
                    mbar_expect_tx(bar_a + 8 * s, TG_A_BYTES);
                    tma_load_2d(sa, &map_a, bar_a + 8 * s, kb * TG_BK, m0);
                    if (leader) mbar_expect_tx(bar_b + 8 * s, 4 * T2_BH_BYTES);  // hi + lo halves of both CTAs
                    const uint32_t bb = mapa_rank0(bar_b + 8 * s);
                    tma_load_2d_2sm(sa + 2 * TG_A_BYTES, &map_bhi, bb, kb * TG_BK, n_half);
                    tma_load_2d_2sm(sa + 2 * TG_A_BYTES + T2_BH_BYTES, &map_blo, bb, kb * TG_BK, n_half);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader) {
            const uint32_t idesc = make_idesc_tf32_m256(TG_BN);
            const int total_kb = u1 - u0;
            for (int kb = 0; kb < total_kb; ++kb) {
                const int s = kb % T2_STAGES, b = kb % TG_NBUF;
                mbar_wait(bar_b + 8 * s, (kb / T2_STAGES) & 1);           // both B halves landed
                mbar_wait(bar_conv + 8 * s, (kb / T2_STAGES) & 1);        // A split done in both CTAs
                mbar_wait(bar_tempty + 8 * b, ((kb / TG_NBUF) & 1) ^ 1);  // accumulator buffer drained in both CTAs
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    if (!(p.dbg & 4)) {
                        const uint32_t sa = smem_u32(smem + (size_t)s * T2_STAGE_BYTES);
                        const uint64_t d_ahi = make_kmajor_sw128_desc(sa);
                        const uint64_t d_alo = make_kmajor_sw128_desc(sa + TG_A_BYTES);
                        const uint64_t d_bhi = make_kmajor_sw128_desc(sa + 2 * TG_A_BYTES);
                        const uint64_t d_blo = make_kmajor_sw128_desc(sa + 2 * TG_A_BYTES + T2_BH_BYTES);
                        const uint32_t d_tmem = tmem_base + (uint32_t)(b * TG_BN);
#pragma unroll
                        for (int ks = 0; ks < TG_BK / 8; ++ks) {
                            const uint64_t ko = (uint64_t)(ks * 32 >> 4);
                            umma_tf32_2sm(d_tmem, d_ahi + ko, d_blo + ko, idesc, ks != 0);
                            umma_tf32_2sm(d_tmem, d_alo + ko, d_bhi + ko, idesc, 1);
                        }
#pragma unroll
                        for (int ks = 0; ks < TG_BK / 8; ++ks) {
                            const uint64_t ko = (uint64_t)(ks * 32 >> 4);
                            umma_tf32_2sm(d_tmem, d_ahi + ko, d_bhi + ko, idesc, 1);
                        }
                    }
                    umma_commit_2sm(bar_empty + 8 * s, 3);   // stage free in both CTAs
                    umma_commit_2sm(bar_tfull + 8 * b, 3);   // chunk complete in both CTAs
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4 && warp < 4 + TG_CONV_WARPS) {
        // ===================== converter warps (both CTAs: own A rows) =====================
        const int ct = threadIdx.x - 128;
        const int total_kb = u1 - u0;
        for (int kb = 0; kb < total_kb; ++kb) {
            const int s = kb % T2_STAGES;
            mbar_wait(bar_a + 8 * s, (kb / T2_STAGES) & 1);
            const uint32_t stage = smem_u32(smem + (size_t)s * T2_STAGE_BYTES);
            constexpr int CA = TG_A_BYTES / 16;
#pragma unroll
            for (int i = 0; i < CA / (32 * TG_CONV_WARPS); ++i) {
                if (p.dbg & 1) break;
                const uint32_t hi = stage + 16u * (uint32_t)(ct + i * 32 * TG_CONV_WARPS);
                const uint32_t lo = hi + (uint32_t)TG_A_BYTES;
                const uint4 v = lds128(hi);
                uint4 h, l;
                split_tf32(__uint_as_float(v.x), h.x, l.x);
                split_tf32(__uint_as_float(v.y), h.y, l.y);
                split_tf32(__uint_as_float(v.z), h.z, l.z);
                split_tf32(__uint_as_float(v.w), h.w, l.w);
                sts128(hi, h);
                sts128(lo, l);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_rank0(bar_conv + 8 * s));
        }
    } else if (warp >= 8) {
        // ===================== drain + epilogue warps (both CTAs: own 128 rows) =====================
        const int q = warp & 3;
        const int half = (warp - 8) >> 2;
        int it = 0;
        const int dt = threadIdx.x - 256;
        for (int hi = u1; hi > u0;) {
        const int tile = (hi - 1) / num_kb;
        const int kb_begin = max(u0 - tile * num_kb, 0), kb_end = hi - tile * num_kb;
        hi = tile * num_kb + kb_begin;
        const bool head_part = kb_end < num_kb;
        const bool tail_part = kb_begin > 0;
        const int m0 = tile_m0(tile), n_base = (tile % tiles_n) * TG_BN;
        const int row = m0 + q * 32 + lane;
        float acc[80];
#pragma unroll
        for (int j = 0; j < 80; ++j) acc[j] = 0.f;
        uint32_t mbits[3] = {0u, 0u, 0u};
        if (MASK && row < p.M && !head_part) {
            const int nb0 = n_base + half * 80;
            const int w0 = nb0 >> 5, sh = nb0 & 31;
            const uint32_t *mr = p.mask + (size_t)row * p.wpr;
            uint32_t w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) w[i] = (w0 + i < p.wpr) ? __ldg(mr + w0 + i) : 0u;
#pragma unroll
            for (int i = 0; i < 3; ++i) mbits[i] = __funnelshift_r(w[i], w[i + 1], sh);
        }
        for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
            const int b = it % TG_NBUF;
            mbar_wait(bar_tfull + 8 * b, (it / TG_NBUF) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * TG_BN + half * 80);
#pragma unroll
            for (int c0 = 0; c0 < 80; c0 += 16) {
                if (p.dbg & 2) break;
                uint32_t r[16];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(taddr + (uint32_t)c0));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[c0 + j] += __uint_as_float(r[j]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_rank0(bar_tempty + 8 * b));
        }
        if (head_part) {
            float4 *slot = reinterpret_cast<float4 *>(p.part) + (size_t)blockIdx.x * (TG_BM * TG_BN / 4);
#pragma unroll
            for (int v = 0; v < 20; ++v)
                __stcg(slot + v * 256 + dt, make_float4(acc[4 * v], acc[4 * v + 1], acc[4 * v + 2], acc[4 * v + 3]));
            __threadfence();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (dt == 0) {
                __threadfence();
                *reinterpret_cast<volatile int *>(p.flags + blockIdx.x) = 1;
            }
            continue;
        }
        if (tail_part) {
            const volatile int *flag = p.flags + blockIdx.x - 2;
            while (*flag == 0) __nanosleep(64);
            __threadfence();
            const float4 *slot = reinterpret_cast<const float4 *>(p.part) + (size_t)(blockIdx.x - 2) * (TG_BM * TG_BN / 4);
#pragma unroll
            for (int v = 0; v < 20; ++v) {
                const float4 t = __ldcg(slot + v * 256 + dt);
                acc[4 * v] = t.x + acc[4 * v]; acc[4 * v + 1] = t.y + acc[4 * v + 1];
                acc[4 * v + 2] = t.z + acc[4 * v + 2]; acc[4 * v + 3] = t.w + acc[4 * v + 3];
            }
        }
        const int nb = n_base + half * 80;
        if (row < p.M) {
            float *dst = p.D + (size_t)row * p.N + nb;
            if (MASK) {
#pragma unroll
                for (int j = 0; j < 80; ++j) acc[j] = (mbits[j >> 5] >> (j & 31)) & 1u ? acc[j] : 0.f;
            }
            if ((p.N & 3) == 0) {
#pragma unroll
                for (int v = 0; v < 20; ++v)
                    if (nb + 4 * v + 4 <= p.N)
                        *reinterpret_cast<float4 *>(dst + 4 * v) =
                            make_float4(acc[4 * v], acc[4 * v + 1], acc[4 * v + 2], acc[4 * v + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < 80; ++j)
                    if (nb + j < p.N) dst[j] = acc[j];
            }
        }
        }  // tile loop
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();  // no CTA leaves (or frees TMEM) while its peer may still signal it or the MMAs still write
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// =================================================================================================
// wgrad:  gW (K_in x N_out) = X^T (K_in x M) . gH (M x N_out)      -- reduction over the M rows.
//
// Both operands are activations that are row-major over the REDUCTION index (X[row][k_in],
// gH[row][n_out]), i.e. "MN-major" for the tensor core.  Instead of MN-major descriptors the
// converter warps -- which rewrite every element anyway to split it -- also transpose: TMA brings
// un-swizzled raw tiles raw[32 rows][128 k_in] / raw[32 rows][160 n_out]; each converter thread reads
// 4 consecutive rows of one column (conflict-free) and writes one 16-byte hi chunk and one lo chunk
// into the canonical K-major SWIZZLE_128B operand tiles (chunk ^= row & 7).  From there on the
// pipeline is the one above: 12 UMMAs per 32-row k-block into one of 3 TMEM buffers, drained and added
// in FP32 registers.  The reduction is split over gridDim.z; partial tiles go to a workspace and are
// summed in a fixed order by splitk_reduce (deterministic).
//   warp 0: TMA   warp 1: MMA   warps 4-11: converter/transposer   warps 12-27: drain (40 columns each)
constexpr int WG_NR = 2;                          // raw ring slots
constexpr int WG_NS = 2;                          // split (operand) ring stages
constexpr int WG_RAW_A = TG_BK * TG_BM * 4;       // 16 KB  raw[32][128]
constexpr int WG_RAW_B = TG_BK * TG_BN * 4;       // 20 KB  raw[32][160]
constexpr int WG_RAW_BYTES = WG_RAW_A + WG_RAW_B;
constexpr int WG_THREADS = 896;
constexpr int WG_CONV_WARPS = 8;
constexpr int WG_DRAIN_WARPS = 16;

struct WGParams {
    int M, Kin, Nout;     // X is M x Kin, gH is M x Nout
    int kb_per_split, num_kb_total;
    float *part;          // (gridDim.z, Kin, Nout)
};

__device__ __forceinline__ void split4(const float v0, const float v1, const float v2, const float v3, uint4 &h,
                                       uint4 &l) {
    split_tf32(v0, h.x, l.x);
    split_tf32(v1, h.y, l.y);
    split_tf32(v2, h.z, l.z);
    split_tf32(v3, h.w, l.w);
}

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tf32x3_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_g,
                    const WGParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *split_base = smem;                                   // WG_NS x TG_STAGE_BYTES (1024-aligned pieces)
    uint8_t *raw_base = smem + (size_t)WG_NS * TG_STAGE_BYTES;    // WG_NR x WG_RAW_BYTES
    __shared__ __align__(8) uint64_t bars[2 * WG_NR + 2 * WG_NS + 2 * TG_NBUF];
    __shared__ uint32_t tmem_base_slot;
    const uint32_t bar_rfull = smem_u32(&bars[0]);
    const uint32_t bar_rempty = smem_u32(&bars[WG_NR]);
    const uint32_t bar_sfull = smem_u32(&bars[2 * WG_NR]);
    const uint32_t bar_sempty = smem_u32(&bars[2 * WG_NR + WG_NS]);
    const uint32_t bar_tfull = smem_u32(&bars[2 * WG_NR + 2 * WG_NS]);
    const uint32_t bar_tempty = smem_u32(&bars[2 * WG_NR + 2 * WG_NS + TG_NBUF]);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * TG_BM;   // k_in offset of this tile
    const int n0 = blockIdx.y * TG_BN;   // n_out offset
    const int kb_begin = blockIdx.z * p.kb_per_split;
    const int nk = min(p.num_kb_total, kb_begin + p.kb_per_split) - kb_begin;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_g) : "memory");
        for (int i = 0; i < WG_NR; ++i) { mbar_init(bar_rfull + 8 * i, 1); mbar_init(bar_rempty + 8 * i, WG_CONV_WARPS); }
        for (int i = 0; i < WG_NS; ++i) { mbar_init(bar_sfull + 8 * i, WG_CONV_WARPS); mbar_init(bar_sempty + 8 * i, 1); }
        for (int i = 0; i < TG_NBUF; ++i) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, WG_DRAIN_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;
    pdl_launch_dependents();
    pdl_wait();  // set-up done; the predecessors' global memory is read from here on

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < nk; ++i) {
                const int r = i % WG_NR;
                mbar_wait(bar_rempty + 8 * r, ((i / WG_NR) & 1) ^ 1);
                const uint32_t ra = smem_u32(raw_base + (size_t)r * WG_RAW_BYTES);
                mbar_expect_tx(bar_rfull + 8 * r, WG_RAW_BYTES);
                tma_load_2d(ra, &map_x, bar_rfull + 8 * r, m0, (kb_begin + i) * TG_BK);
                tma_load_2d(ra + WG_RAW_A, &map_g, bar_rfull + 8 * r, n0, (kb_begin + i) * TG_BK);
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = make_idesc_tf32(TG_BN);
        for (int i = 0; i < nk; ++i) {
            const int s = i % WG_NS, b = i % TG_NBUF;
            mbar_wait(bar_sfull + 8 * s, (i / WG_NS) & 1);
            mbar_wait(bar_tempty + 8 * b, ((i / TG_NBUF) & 1) ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint32_t sa = smem_u32(split_base + (size_t)s * TG_STAGE_BYTES);
                const uint64_t d_ahi = make_kmajor_sw128_desc(sa);
                const uint64_t d_alo = make_kmajor_sw128_desc(sa + TG_A_BYTES);
                const uint64_t d_bhi = make_kmajor_sw128_desc(sa + 2 * TG_A_BYTES);
                const uint64_t d_blo = make_kmajor_sw128_desc(sa + 2 * TG_A_BYTES + TG_B_BYTES);
                const uint32_t d_tmem = tmem_base + (uint32_t)(b * TG_BN);
#pragma unroll
                for (int ks = 0; ks < TG_BK / 8; ++ks) {
                    const uint64_t ko = (uint64_t)(ks * 32 >> 4);
                    umma_tf32(d_tmem, d_ahi + ko, d_blo + ko, idesc, ks != 0);
                    umma_tf32(d_tmem, d_alo + ko, d_bhi + ko, idesc, 1);
                }
#pragma unroll
                for (int ks = 0; ks < TG_BK / 8; ++ks) {
                    const uint64_t ko = (uint64_t)(ks * 32 >> 4);
                    umma_tf32(d_tmem, d_ahi + ko, d_bhi + ko, idesc, 1);
                }
                umma_commit(bar_sempty + 8 * s);
                umma_commit(bar_tfull + 8 * b);
            }
            __syncwarp();
        }
    } else if (warp >= 4 && warp < 4 + WG_CONV_WARPS) {
        // ===================== converter + transposer =====================
        const int ct = threadIdx.x - 128;  // 0..255
        for (int i = 0; i < nk; ++i) {
            const int r = i % WG_NR, s = i % WG_NS;
            mbar_wait(bar_rfull + 8 * r, (i / WG_NR) & 1);
            mbar_wait(bar_sempty + 8 * s, ((i / WG_NS) & 1) ^ 1);
            const float *rawA = reinterpret_cast<const float *>(raw_base + (size_t)r * WG_RAW_BYTES);
            const float *rawB = rawA + TG_BK * TG_BM;
            uint8_t *st = split_base + (size_t)s * TG_STAGE_BYTES;
#pragma unroll
            for (int j = 0; j < (TG_BM * 8) / 256; ++j) {        // A: 128 columns x 8 row-quads
                const int t = ct + j * 256;
                const int col = t & (TG_BM - 1), q = t / TG_BM;
                uint4 h, l;
                split4(rawA[(4 * q + 0) * TG_BM + col], rawA[(4 * q + 1) * TG_BM + col],
                       rawA[(4 * q + 2) * TG_BM + col], rawA[(4 * q + 3) * TG_BM + col], h, l);
                const int off = col * 128 + ((q ^ (col & 7)) << 4);
                *reinterpret_cast<uint4 *>(st + off) = h;
                *reinterpret_cast<uint4 *>(st + TG_A_BYTES + off) = l;
            }
#pragma unroll
            for (int j = 0; j < (TG_BN * 8) / 256; ++j) {        // B: 160 columns x 8 row-quads
                const int t = ct + j * 256;
                const int col = t % TG_BN, q = t / TG_BN;
                uint4 h, l;
                split4(rawB[(4 * q + 0) * TG_BN + col], rawB[(4 * q + 1) * TG_BN + col],
                       rawB[(4 * q + 2) * TG_BN + col], rawB[(4 * q + 3) * TG_BN + col], h, l);
                const int off = col * 128 + ((q ^ (col & 7)) << 4);
                *reinterpret_cast<uint4 *>(st + 2 * TG_A_BYTES + off) = h;
                *reinterpret_cast<uint4 *>(st + 2 * TG_A_BYTES + TG_B_BYTES + off) = l;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(bar_sfull + 8 * s);
                mbar_arrive(bar_rempty + 8 * r);
            }
        }
    } else if (warp >= 4 + WG_CONV_WARPS) {
        // ===================== drain + epilogue =====================
        const int q = warp & 3;
        const int cpart = (warp - 4 - WG_CONV_WARPS) >> 2;   // 0..3 -> columns [40*cpart, 40*cpart + 40)
        float acc[40];
#pragma unroll
        for (int j = 0; j < 40; ++j) acc[j] = 0.f;
        for (int i = 0; i < nk; ++i) {
            const int b = i % TG_NBUF;
            mbar_wait(bar_tfull + 8 * b, (i / TG_NBUF) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * TG_BN + cpart * 40);
            uint32_t r0[16], r1[16], r2[8];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                : "=r"(r0[0]), "=r"(r0[1]), "=r"(r0[2]), "=r"(r0[3]), "=r"(r0[4]), "=r"(r0[5]), "=r"(r0[6]), "=r"(r0[7]),
                  "=r"(r0[8]), "=r"(r0[9]), "=r"(r0[10]), "=r"(r0[11]), "=r"(r0[12]), "=r"(r0[13]), "=r"(r0[14]), "=r"(r0[15])
                : "r"(taddr));
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                : "=r"(r1[0]), "=r"(r1[1]), "=r"(r1[2]), "=r"(r1[3]), "=r"(r1[4]), "=r"(r1[5]), "=r"(r1[6]), "=r"(r1[7]),
                  "=r"(r1[8]), "=r"(r1[9]), "=r"(r1[10]), "=r"(r1[11]), "=r"(r1[12]), "=r"(r1[13]), "=r"(r1[14]), "=r"(r1[15])
                : "r"(taddr + 16u));
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                : "=r"(r2[0]), "=r"(r2[1]), "=r"(r2[2]), "=r"(r2[3]), "=r"(r2[4]), "=r"(r2[5]), "=r"(r2[6]), "=r"(r2[7])
                : "r"(taddr + 32u));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; ++j) { acc[j] += __uint_as_float(r0[j]); acc[16 + j] += __uint_as_float(r1[j]); }
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[32 + j] += __uint_as_float(r2[j]);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * b);
        }
        const int m = m0 + q * 32 + lane;            // k_in index (row of gW)
        const int nb = n0 + cpart * 40;
        if (m < p.Kin) {
            float *dst = p.part + ((size_t)blockIdx.z * p.Kin + m) * p.Nout + nb;
#pragma unroll
            for (int v = 0; v < 10; ++v)
                if (nb + 4 * v + 4 <= p.Nout)
                    *reinterpret_cast<float4 *>(dst + 4 * v) =
                        make_float4(acc[4 * v], acc[4 * v + 1], acc[4 * v + 2], acc[4 * v + 3]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// Split a weight matrix into TF32 hi/lo parts, optionally transposing:
//   transpose = 0: out[r, c] = split(src[r, c])           (rows x cols) -> (rows x cols)
//   transpose = 1: out[c, r] = split(src[r, c])           (rows x cols) -> (cols x rows)
__global__ void split_tf32_kernel(const float *__restrict__ src, int rows, int cols, int transpose,
                                  float *__restrict__ hi, float *__restrict__ lo, int *__restrict__ flags,
                                  int n_flags) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_flags) flags[i] = 0;  // the GEMM's partial-tile flags (same stream, launched right after)
    if (i >= (long long)rows * cols) return;
    const int r = (int)(i / cols), c = (int)(i % cols);
    uint32_t h, l;
    split_tf32(src[i], h, l);
    const size_t o = transpose ? (size_t)c * rows + r : (size_t)i;
    hi[o] = __uint_as_float(h);
    lo[o] = __uint_as_float(l);
}

// The weight splits of a whole GCN pass in one launch (csrc/gcn_stack.cu): blockIdx.y = layer.  Also zeroes the
// partial-tile flags of every layer's GEMM (layer l owns flags[l * flags_per_layer ...]).
struct SplitBatch {
    const float *src[TG_SPLIT_BATCH];
    float *hi[TG_SPLIT_BATCH];
    float *lo[TG_SPLIT_BATCH];
    int rows[TG_SPLIT_BATCH], cols[TG_SPLIT_BATCH];
    int transpose;
    int *flags;
    int n_flags;
};
__global__ void split_tf32_batched_kernel(const __grid_constant__ SplitBatch b) {
    pdl_launch_dependents();
    pdl_wait();
    const int layer = blockIdx.y;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (layer == 0 && i < b.n_flags) b.flags[i] = 0;
    const int rows = b.rows[layer], cols = b.cols[layer];
    if (i >= (long long)rows * cols) return;
    const int r = (int)(i / cols), c = (int)(i % cols);
    uint32_t h, l;
    split_tf32(b.src[layer][i], h, l);
    const size_t o = b.transpose ? (size_t)c * rows + r : (size_t)i;
    b.hi[layer][o] = __uint_as_float(h);
    b.lo[layer][o] = __uint_as_float(l);
}

// ------------------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 2-D fp32 row-major (rows x cols, cols contiguous) tensor map.
//   swizzled : box = (32 cols, box_rows), SWIZZLE_128B    (K-major operand tiles)
//   plain    : box = (box_cols, box_rows), no swizzle      (raw tiles for the transposing converter)
static int make_map_ex(CUtensorMap *map, const float *base, int64_t rows, int64_t cols, int box_cols, int box_rows,
                       bool swizzle);
static int make_map(CUtensorMap *map, const float *base, int64_t rows, int64_t cols, int box_rows) {
    return make_map_ex(map, base, rows, cols, TG_BK, box_rows, true);
}
static int make_map_ex(CUtensorMap *map, const float *base, int64_t rows, int64_t cols, int box_cols, int box_rows,
                       bool swizzle) {
    EncodeTiledFn fn = encode_fn();
    PTK_REQUIRE(fn != nullptr, PTK_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable in this driver");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PTK_REQUIRE(r == CUDA_SUCCESS, PTK_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld box_rows=%d",
                (int)r, (long long)rows, (long long)cols, box_rows);
    return PTK_OK;
}

// for the other translation units (gcn_linear.cu: TMA-staged exact forward)
int make_tensor_map_2d(CUtensorMap *map, const float *base, int64_t rows, int64_t cols, int box_cols, int box_rows,
                       bool swizzle) {
    return make_map_ex(map, base, rows, cols, box_cols, box_rows, swizzle);
}

// bits[row * wpr + j] bit i = act[row, 32 j + i] > 0 (0 beyond N).  One warp per row, coalesced.
__global__ void __launch_bounds__(256)
relu_bits_kernel(const float *__restrict__ act, long long M, int N, int wpr, uint32_t *__restrict__ bits) {
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const float *a = act + (size_t)row * N;
    uint32_t mine = 0u;
    for (int j = 0; j < wpr; ++j) {
        const int c = 32 * j + lane;
        const uint32_t word = __ballot_sync(0xffffffffu, c < N && a[c] > 0.f);
        if ((j & 31) == lane) mine = word;
        if ((j & 31) == 31 || j == wpr - 1) {  // flush up to 32 words with one coalesced store
            const int j0 = j & ~31;
            if (j0 + lane <= j) bits[(size_t)row * wpr + j0 + lane] = mine;
        }
    }
}

// cudaFuncAttributeMaxDynamicSharedMemorySize has to be set once per DEVICE (a process may drive several):
// one bit per (kernel family, device ordinal).
static unsigned long long g_smem_optin[2] = {0ull, 0ull};
static bool smem_optin_done(int family) {
    int dev = 0;
    return cudaGetDevice(&dev) == cudaSuccess && dev < 64 && ((g_smem_optin[family] >> dev) & 1ull);
}
static void smem_optin_mark(int family) {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && dev < 64) g_smem_optin[family] |= 1ull << dev;
}

bool tf32x3_eligible(const void *A, const void *D, int64_t M, int64_t K, int64_t N) {
    // TMA needs 16-byte aligned bases and row pitches; tiny reductions / outputs stay on the SIMT path
    return M >= 1 && K >= 32 && N >= 16 && (K % 4) == 0 && (((uintptr_t)A) % 16) == 0 && (((uintptr_t)D) % 16) == 0;
}

// split B (hi, lo) + the packed ReLU mask of D (M rows x ceil(N/32) words; D has max(K, N) columns at most:
// the same function sizes forward (D = M x N) and dgrad (D = M x K) workspaces)
static size_t tf32x3_part_bytes() { return (size_t)sm_count() * (TG_BM * TG_BN * 4 + 4) + 512; }
size_t tf32x3_workspace_bytes(int64_t M, int64_t K, int64_t N) {
    const int64_t cols = K > N ? K : N;
    return 2 * sizeof(float) * (size_t)K * (size_t)N + 512 + sizeof(uint32_t) * (size_t)M * (size_t)ceil_div(cols, 32) +
           tf32x3_part_bytes();
}

// PTK_TG_PAIR=1 selects the 2-CTA cluster form (weight tiles multicast over the pair).  Measured on B200 at M = 31184,
// K = N = 300 (tools/tc_gemm_probe.py): the operand floor (TMA only: split, MMAs and drain disabled) falls from 48.5 to
// 37.1 us, the kernel itself stays at 53.3 us (53.6 us single) -- with 3 stages of 72 KB it is bound by the depth of the
// TMA -> split -> MMA -> release ring, not by L2 reads any more (MMAs disabled: 40.0 us).  Left off: no gain without a
// deeper ring; the cta_group::2 form below (PTK_TG_PAIR=3) has that ring but pays for it in cross-CTA barrier hops.
static bool tg_pair_mode() { return PTK_TUNING_ENV("PTK_TG_PAIR") == 1 || PTK_TUNING_ENV("PTK_TG_PAIR") == 3; }
static bool tg_2sm_mode() { return PTK_TUNING_ENV("PTK_TG_PAIR") == 3; }  // cta_group::2 MMAs (gemm_tf32x3_2sm_kernel)

// CTAs to launch: one per SM at most; PAIR: whole clusters of two, one per 256-row pair x column tile at most
static int tg_num_ctas(int64_t M, int64_t N) {
    const int64_t tiles_m = ceil_div(M, TG_BM), tiles_n = ceil_div(N, TG_BN);
    if (tg_pair_mode()) {
        const int64_t ptiles = ceil_div(tiles_m, 2) * tiles_n;
        const int64_t clusters = ptiles < sm_count() / 2 ? ptiles : sm_count() / 2;
        return (int)(2 * clusters);
    }
    const int64_t tiles = tiles_m * tiles_n;
    return (int)(tiles < sm_count() ? tiles : sm_count());
}

template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl_cluster2(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                       Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static int tg_launch(const float *A, const float *b_hi, const float *b_lo, const uint32_t *mask, int64_t M, int64_t K,
                     int64_t N, float *D, float *part, int *flags, int n_ctas, cudaStream_t st) {
    const bool pair = tg_pair_mode();
    CUtensorMap map_a, map_bhi, map_blo;
    int rc = make_map(&map_a, A, M, K, TG_BM);
    if (rc) return rc;
    rc = make_map(&map_bhi, b_hi, N, K, pair ? TG_BN / 2 : TG_BN);  // (N x K) K-major; rows beyond N are zero-filled
    if (rc) return rc;
    rc = make_map(&map_blo, b_lo, N, K, pair ? TG_BN / 2 : TG_BN);
    if (rc) return rc;

    TGParams p;
    p.M = (int)M; p.N = (int)N; p.K = (int)K; p.D = D; p.mask = mask; p.wpr = (int)ceil_div(N, 32);
    static int dbg = -1;
    if (dbg < 0) { const char *e = getenv("PTK_TG_DEBUG"); dbg = e ? atoi(e) : 0; }
    p.dbg = dbg;
    p.part = part;
    p.flags = flags;
    const size_t smem = (size_t)TG_STAGES * TG_STAGE_BYTES + 1024;
    dim3 grid((unsigned)n_ctas);
    if (tg_2sm_mode()) {
        const size_t smem2 = (size_t)T2_STAGES * T2_STAGE_BYTES + 1024;
        static unsigned long long optin2 = 0ull;
        int dev = 0;
        PTK_CHECK_CUDA(cudaGetDevice(&dev));
        if (dev >= 64 || !((optin2 >> dev) & 1ull)) {
            PTK_CHECK_CUDA(cudaFuncSetAttribute(gemm_tf32x3_2sm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            PTK_CHECK_CUDA(cudaFuncSetAttribute(gemm_tf32x3_2sm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            if (dev < 64) optin2 |= 1ull << dev;
        }
        const cudaError_t le2 =
            mask ? launch_pdl_cluster2(gemm_tf32x3_2sm_kernel<true>, grid, dim3(TG_THREADS), smem2, st, map_a, map_bhi, map_blo, p)
                 : launch_pdl_cluster2(gemm_tf32x3_2sm_kernel<false>, grid, dim3(TG_THREADS), smem2, st, map_a, map_bhi, map_blo, p);
        PTK_CHECK_CUDA(le2);
        PTK_CHECK_LAUNCH();
        return PTK_OK;
    }
    if (!smem_optin_done(0)) {  // the attribute is per device
        PTK_CHECK_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PTK_CHECK_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PTK_CHECK_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PTK_CHECK_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_optin_mark(0);
    }
    cudaError_t le;
    if (pair)
        le = mask ? launch_pdl_cluster2(gemm_tf32x3_kernel<true, true>, grid, dim3(TG_THREADS), smem, st, map_a, map_bhi, map_blo, p)
                  : launch_pdl_cluster2(gemm_tf32x3_kernel<false, true>, grid, dim3(TG_THREADS), smem, st, map_a, map_bhi, map_blo, p);
    else
        le = mask ? launch_pdl(gemm_tf32x3_kernel<true, false>, grid, dim3(TG_THREADS), smem, st, map_a, map_bhi, map_blo, p)
                  : launch_pdl(gemm_tf32x3_kernel<false, false>, grid, dim3(TG_THREADS), smem, st, map_a, map_bhi, map_blo, p);
    PTK_CHECK_CUDA(le);
    PTK_CHECK_LAUNCH();
    return PTK_OK;
}

// D (M x N) = A (M x K) . Bsrc, where Bsrc is either (K x N) row-major [b_is_kn = 1: transposed during the
// split] or (N x K) row-major [b_is_kn = 0].  act (optional): D masked by act > 0.
int gemm_tf32x3(const float *A, const float *Bsrc, int b_is_kn, const float *act, int64_t M, int64_t K, int64_t N,
                float *D, void *workspace, size_t workspace_bytes, cudaStream_t st, const uint32_t *act_bits) {
    PTK_REQUIRE(workspace && workspace_bytes >= tf32x3_workspace_bytes(M, K, N), PTK_ERR_WORKSPACE,
                "gemm_tf32x3: workspace too small");
    float *b_hi = reinterpret_cast<float *>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    float *b_lo = b_hi + (size_t)K * N;
    uint32_t *mask = reinterpret_cast<uint32_t *>(((uintptr_t)(b_lo + (size_t)K * N) + 255) & ~(uintptr_t)255);
    const int wpr = (int)ceil_div(N, 32);
    // partial tiles + flags behind the (always reserved) mask area
    float *part = reinterpret_cast<float *>(
        ((uintptr_t)(mask + (size_t)M * (size_t)ceil_div(K > N ? K : N, 32)) + 255) & ~(uintptr_t)255);
    int *flags = reinterpret_cast<int *>(part + (size_t)sm_count() * (TG_BM * TG_BN));
    const int n_ctas = tg_num_ctas(M, N);
    if (act && act_bits) {
        mask = const_cast<uint32_t *>(act_bits);  // packed by the forward GEMM that consumed `act` (same layout)
    } else if (act) {
        relu_bits_kernel<<<(unsigned)ceil_div(M, 8), 256, 0, st>>>(act, (long long)M, (int)N, wpr, mask);
        PTK_CHECK_LAUNCH();
    }
    const long long elems = (long long)K * N;
    if (b_is_kn)
        split_tf32_kernel<<<(unsigned)ceil_div(elems, 256), 256, 0, st>>>(Bsrc, (int)K, (int)N, 1, b_hi, b_lo, flags, n_ctas);
    else
        split_tf32_kernel<<<(unsigned)ceil_div(elems, 256), 256, 0, st>>>(Bsrc, (int)N, (int)K, 0, b_hi, b_lo, flags, n_ctas);
    PTK_CHECK_LAUNCH();

    return tg_launch(A, b_hi, b_lo, act ? mask : nullptr, M, K, N, D, part, flags, n_ctas, st);
}

// One GEMM whose B operand was split before (tf32x3_presplit_batched); mask_ws (M x ceil(N/32) words) is only used
// when `act` comes without its packed mask.  flags: this GEMM's own n_ctas ints, zeroed by the batched split.
int gemm_tf32x3_presplit(const float *A, const float *b_hi, const float *b_lo, const float *act, const uint32_t *act_bits,
                         int64_t M, int64_t K, int64_t N, float *D, uint32_t *mask_ws, float *part, int *flags,
                         cudaStream_t st) {
    const int n_ctas = tg_num_ctas(M, N);
    const uint32_t *mask = nullptr;
    if (act && act_bits) {
        mask = act_bits;
    } else if (act) {
        PTK_REQUIRE(mask_ws, PTK_ERR_WORKSPACE, "gemm_tf32x3_presplit: no mask workspace");
        relu_bits_kernel<<<(unsigned)ceil_div(M, 8), 256, 0, st>>>(act, (long long)M, (int)N, (int)ceil_div(N, 32), mask_ws);
        PTK_CHECK_LAUNCH();
        mask = mask_ws;
    }
    return tg_launch(A, b_hi, b_lo, mask, M, K, N, D, part, flags, n_ctas, st);
}

size_t tf32x3_part_floats() { return (size_t)sm_count() * (TG_BM * TG_BN); }
int tf32x3_max_ctas() { return sm_count(); }

// Split n weight matrices (src[i]: rows[i] x cols[i] row-major) into hi / lo in one launch per 32 layers and zero
// `n_flags` partial-tile flags.  transpose: write (cols x rows) -- the forward's (K x N) weights as K-major B.
int tf32x3_presplit_batched(int n, const float *const *src, const int *rows, const int *cols, int transpose,
                            float *const *hi, float *const *lo, int *flags, int n_flags, cudaStream_t st) {
    for (int base = 0; base < n; base += TG_SPLIT_BATCH) {
        SplitBatch b;
        const int cnt = n - base < TG_SPLIT_BATCH ? n - base : TG_SPLIT_BATCH;
        long long max_elems = base == 0 ? n_flags : 0;
        for (int i = 0; i < TG_SPLIT_BATCH; ++i) {
            const int j = i < cnt ? base + i : base;
            b.src[i] = src[j]; b.hi[i] = hi[j]; b.lo[i] = lo[j]; b.rows[i] = rows[j]; b.cols[i] = cols[j];
            const long long e = (long long)rows[j] * cols[j];
            if (e > max_elems) max_elems = e;
        }
        b.transpose = transpose;
        b.flags = flags;
        b.n_flags = base == 0 ? n_flags : 0;
        if (max_elems <= 0) continue;
        launch_pdl(split_tf32_batched_kernel, dim3((unsigned)ceil_div(max_elems, 256), (unsigned)cnt), dim3(256), 0, st, b);
        PTK_CHECK_LAUNCH();
    }
    return PTK_OK;
}

// ---- wgrad host side
struct WGPlan { int tiles_m, tiles_n, splits, kb_per_split, num_kb; };
static WGPlan plan_wgrad(int64_t M, int64_t Kin, int64_t Nout) {
    WGPlan w;
    w.tiles_m = (int)ceil_div(Kin, TG_BM);
    w.tiles_n = (int)ceil_div(Nout, TG_BN);
    w.num_kb = (int)ceil_div(M, TG_BK);
    int want = sm_count() / (w.tiles_m * w.tiles_n);
    if (want < 1) want = 1;
    int max_s = w.num_kb / 4 > 0 ? w.num_kb / 4 : 1;
    if (want > max_s) want = max_s;
    w.kb_per_split = (int)ceil_div(w.num_kb, want);
    w.splits = (int)ceil_div(w.num_kb, w.kb_per_split);
    return w;
}

bool wgrad_tf32x3_eligible(const void *X, const void *gH, int64_t M, int64_t Kin, int64_t Nout) {
    return M >= 32 && Kin >= 32 && Nout >= 16 && (Kin % 4) == 0 && (Nout % 4) == 0 && (((uintptr_t)X) % 16) == 0 &&
           (((uintptr_t)gH) % 16) == 0;
}

size_t wgrad_tf32x3_workspace_bytes(int64_t M, int64_t Kin, int64_t Nout) {
    const WGPlan w = plan_wgrad(M, Kin, Nout);
    return sizeof(float) * (size_t)w.splits * (size_t)Kin * (size_t)Nout + 256;
}

// part (splits, Kin, Nout) in `workspace`; returns the number of splits through *n_splits
int wgrad_tf32x3(const float *X, const float *gH, int64_t M, int64_t Kin, int64_t Nout, void *workspace,
                 size_t workspace_bytes, float **part_out, int *n_splits, cudaStream_t st) {
    PTK_REQUIRE(workspace && workspace_bytes >= wgrad_tf32x3_workspace_bytes(M, Kin, Nout), PTK_ERR_WORKSPACE,
                "wgrad_tf32x3: workspace too small");
    const WGPlan w = plan_wgrad(M, Kin, Nout);
    float *part = reinterpret_cast<float *>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    CUtensorMap map_x, map_g;
    int rc = make_map_ex(&map_x, X, M, Kin, TG_BM, TG_BK, false);
    if (rc) return rc;
    rc = make_map_ex(&map_g, gH, M, Nout, TG_BN, TG_BK, false);
    if (rc) return rc;
    WGParams p;
    p.M = (int)M; p.Kin = (int)Kin; p.Nout = (int)Nout; p.kb_per_split = w.kb_per_split; p.num_kb_total = w.num_kb;
    p.part = part;
    const size_t smem = (size_t)WG_NS * TG_STAGE_BYTES + (size_t)WG_NR * WG_RAW_BYTES + 1024;
    if (!smem_optin_done(1)) {
        PTK_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_optin_mark(1);
    }
    dim3 grid((unsigned)w.tiles_m, (unsigned)w.tiles_n, (unsigned)w.splits);
    launch_pdl(wgrad_tf32x3_kernel, grid, dim3(WG_THREADS), smem, st, map_x, map_g, p);
    PTK_CHECK_LAUNCH();
    *part_out = part;
    *n_splits = w.splits;
    return PTK_OK;
}

}  // namespace ptk
