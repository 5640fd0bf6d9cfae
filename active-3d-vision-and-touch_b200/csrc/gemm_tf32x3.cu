// GCN per-vertex linear layers on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// Replaces `torch.matmul(features, self.weight)` (pterotactyl/reconstruction/vision/model.py:352) and
// its dgrad where the layer is a true GEMM (300x300, 448x300; K >= 32).  The parity contract is FP32
// (1e-5 relative), so a single TF32 pass (10-bit mantissa) is not admissible; this kernel runs the
// error-compensated 3xTF32 scheme:
//       a = a_hi + a_lo,  a_hi = tf32_rna(a),  a_lo = tf32_rna(a - a_hi)     (same for b)
//       D = a_hi*b_lo + a_lo*b_hi + a_hi*b_hi        (a_lo*b_lo ~ 2^-22 relative is dropped)
// with FP32 accumulation in tensor memory -- ~22+ mantissa bits per product, i.e. the error level of
// an FP32 FMA chain of the same length, at 1/3 of the TF32 tensor rate instead of the FP32 SIMT rate.
//
//   D (M x N) = A (M x K, row-major: K contiguous)  .  B^T   with B stored (N x K), K contiguous
//     fwd   : A = X,  B = W^T (pre-transposed + pre-split, tiny)          D = H
//     dgrad : A = gH, B = W   (as stored: (K_in x N_out) is "N x K")      D = gX  [* (act > 0)]
//
// One CTA = one 128-row tile x up to 320 columns (two UMMA N-parts), 192 threads:
//   warp 0      TMA producer: raw FP32 A tile (128 x 32) + pre-split B_hi/B_lo tiles per k-block,
//               SWIZZLE_128B, mbarrier expect_tx
//   warp 1      MMA issuer (one elected lane): 3 x tcgen05.mma.kind::tf32 per 8-wide k-step and N-part,
//               accumulators in TMEM; tcgen05.commit releases the smem stage / signals the epilogue
//   warps 2-5   converter: split the raw A tile in place into a_hi (overwrites raw) and a_lo
//               (elementwise, position preserving => swizzle-agnostic), fence.proxy.async, arrive;
//               after the main loop the same four warps are the epilogue: tcgen05.ld 32x32b,
//               optional ReLU-mask, 64-byte vector stores.
#include <cuda.h>

#include "ptk_common.cuh"

namespace ptk {

constexpr int TG_BM = 128;        // rows per CTA (UMMA M, cta_group::1)
constexpr int TG_BK = 32;         // fp32 elements per k-block = 128 B = one SWIZZLE_128B atom row
constexpr int TG_STAGES = 2;
constexpr int TG_THREADS = 192;
constexpr int TG_MAX_BN = 320;    // columns per CTA (<= 2 UMMA parts of <= 256, TMEM has 512 columns)

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
__device__ __forceinline__ uint32_t tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
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

struct TGParams {
    int M, N, K;          // D is M x N, reduction K
    int n0_cols, n1_cols; // UMMA N of part 0 / part 1 (multiples of 16, n1 may be 0); per-CTA BN = n0 + n1
    float *D;
    const float *act;     // optional ReLU mask source, same shape as D
};

template <bool MASK>
__global__ void __launch_bounds__(TG_THREADS, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_bhi,
                   const __grid_constant__ CUtensorMap map_blo, const __grid_constant__ CUtensorMap map_bhi1,
                   const __grid_constant__ CUtensorMap map_blo1, const TGParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: per stage [A_hi(raw) 16 KB][A_lo 16 KB][B_hi bn*128][B_lo bn*128]; 1024-B aligned pieces
    const int bn = p.n0_cols + p.n1_cols;
    const uint32_t a_bytes = TG_BM * TG_BK * 4;            // 16384
    const uint32_t b_bytes = (uint32_t)bn * TG_BK * 4;     // bn * 128 (bn multiple of 16 -> multiple of 2048)
    const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bars[3 * TG_STAGES + 1];
    __shared__ uint32_t tmem_base_slot;
    const uint32_t bar_full = smem_u32(&bars[0]);               // TMA landed        (count 1 + tx)
    const uint32_t bar_conv = smem_u32(&bars[TG_STAGES]);       // A split finished  (count 4: one per warp)
    const uint32_t bar_empty = smem_u32(&bars[2 * TG_STAGES]);  // MMAs done reading (count 1, tcgen05.commit)
    const uint32_t bar_accum = smem_u32(&bars[3 * TG_STAGES]);  // accumulators complete

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * TG_BM;
    const int n_base = blockIdx.y * bn;
    const int num_kb = (p.K + TG_BK - 1) / TG_BK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_bhi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_blo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_bhi1) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_blo1) : "memory");
        for (int s = 0; s < TG_STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_conv + 8 * s, 4);
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: 512 columns (the whole SM's tensor memory; 1 CTA per SM by shared-memory size)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % TG_STAGES;
                const uint32_t ph = (kb / TG_STAGES) & 1;
                mbar_wait(bar_empty + 8 * s, ph ^ 1);
                const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
                const uint32_t sbh = sa + 2 * a_bytes, sbl = sbh + b_bytes;
                mbar_expect_tx(bar_full + 8 * s, a_bytes + 2 * b_bytes);
                tma_load_2d(sa, &map_a, bar_full + 8 * s, kb * TG_BK, m0);
                tma_load_2d(sbh, &map_bhi, bar_full + 8 * s, kb * TG_BK, n_base);
                tma_load_2d(sbl, &map_blo, bar_full + 8 * s, kb * TG_BK, n_base);
                if (p.n1_cols > 0) {
                    tma_load_2d(sbh + p.n0_cols * 128, &map_bhi1, bar_full + 8 * s, kb * TG_BK, n_base + p.n0_cols);
                    tma_load_2d(sbl + p.n0_cols * 128, &map_blo1, bar_full + 8 * s, kb * TG_BK, n_base + p.n0_cols);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc0 = make_idesc_tf32(p.n0_cols);
        const uint32_t idesc1 = make_idesc_tf32(p.n1_cols > 0 ? p.n1_cols : 16);
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % TG_STAGES;
            const uint32_t ph = (kb / TG_STAGES) & 1;
            mbar_wait(bar_full + 8 * s, ph);   // B tiles (and raw A) landed
            mbar_wait(bar_conv + 8 * s, ph);   // A split done
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
                const uint64_t d_ahi = make_kmajor_sw128_desc(sa);
                const uint64_t d_alo = make_kmajor_sw128_desc(sa + a_bytes);
                const uint64_t d_bhi = make_kmajor_sw128_desc(sa + 2 * a_bytes);
                const uint64_t d_blo = make_kmajor_sw128_desc(sa + 2 * a_bytes + b_bytes);
#pragma unroll
                for (int ks = 0; ks < TG_BK / 8; ++ks) {
                    const uint64_t ko = (uint64_t)(ks * 32 >> 4);  // +32 B per k-step inside the swizzle atom
                    const uint32_t acc = (kb | ks) != 0;
                    // part 0
                    umma_tf32(tmem_base, d_ahi + ko, d_blo + ko, idesc0, acc);
                    umma_tf32(tmem_base, d_alo + ko, d_bhi + ko, idesc0, 1);
                    umma_tf32(tmem_base, d_ahi + ko, d_bhi + ko, idesc0, 1);
                    if (p.n1_cols > 0) {
                        const uint64_t bo = (uint64_t)((p.n0_cols * 128) >> 4);
                        umma_tf32(tmem_base + p.n0_cols, d_ahi + ko, d_blo + bo + ko, idesc1, acc);
                        umma_tf32(tmem_base + p.n0_cols, d_alo + ko, d_bhi + bo + ko, idesc1, 1);
                        umma_tf32(tmem_base + p.n0_cols, d_ahi + ko, d_bhi + bo + ko, idesc1, 1);
                    }
                }
                umma_commit(bar_empty + 8 * s);            // stage free once these MMAs have read smem
                if (kb == num_kb - 1) umma_commit(bar_accum);
            }
            __syncwarp();
        }
    } else {
        // ===================== converter warps (2..5) =====================
        const int ct = threadIdx.x - 64;  // 0..127
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % TG_STAGES;
            const uint32_t ph = (kb / TG_STAGES) & 1;
            mbar_wait(bar_full + 8 * s, ph);
            float4 *hi = reinterpret_cast<float4 *>(smem + (size_t)s * stage_bytes);
            float4 *lo = reinterpret_cast<float4 *>(smem + (size_t)s * stage_bytes + a_bytes);
#pragma unroll
            for (int i = 0; i < (TG_BM * TG_BK / 4) / 128; ++i) {
                const int c = ct + i * 128;
                float4 v = hi[c];
                uint4 h, l;
                h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
                l.x = tf32_rna(v.x - __uint_as_float(h.x)); l.y = tf32_rna(v.y - __uint_as_float(h.y));
                l.z = tf32_rna(v.z - __uint_as_float(h.z)); l.w = tf32_rna(v.w - __uint_as_float(h.w));
                reinterpret_cast<uint4 *>(hi)[c] = h;
                reinterpret_cast<uint4 *>(lo)[c] = l;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> async proxy (UMMA)
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_conv + 8 * s);
        }
        // ===================== epilogue (same warps) =====================
        mbar_wait(bar_accum, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;                    // TMEM lane quarter this warp may access
        const int row = m0 + q * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        for (int c0 = 0; c0 < bn; c0 += 16) {
            uint32_t r[16];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(taddr + (uint32_t)c0));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int n = n_base + c0;
            if (row < p.M && n < p.N) {
                float *dst = p.D + (size_t)row * p.N + n;
                if (n + 16 <= p.N && (p.N & 3) == 0) {
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        float4 o = make_float4(__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]),
                                               __uint_as_float(r[4 * v + 2]), __uint_as_float(r[4 * v + 3]));
                        if (MASK) {
                            const float4 a = *reinterpret_cast<const float4 *>(p.act + (size_t)row * p.N + n + 4 * v);
                            o.x = a.x > 0.f ? o.x : 0.f; o.y = a.y > 0.f ? o.y : 0.f;
                            o.z = a.z > 0.f ? o.z : 0.f; o.w = a.w > 0.f ? o.w : 0.f;
                        }
                        *reinterpret_cast<float4 *>(dst + 4 * v) = o;
                    }
                } else {
                    for (int j = 0; j < 16 && n + j < p.N; ++j) {
                        float o = __uint_as_float(r[j]);
                        if (MASK) o = p.act[(size_t)row * p.N + n + j] > 0.f ? o : 0.f;
                        dst[j] = o;
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
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
                                  float *__restrict__ hi, float *__restrict__ lo) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)rows * cols) return;
    const int r = (int)(i / cols), c = (int)(i % cols);
    const float v = src[i];
    const uint32_t h = tf32_rna(v);
    const uint32_t l = tf32_rna(v - __uint_as_float(h));
    const size_t o = transpose ? (size_t)c * rows + r : (size_t)i;
    hi[o] = __uint_as_float(h);
    lo[o] = __uint_as_float(l);
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

// 2-D fp32 row-major (rows x cols, cols contiguous) tensor map, box = (32 cols, box_rows), SWIZZLE_128B
static int make_map(CUtensorMap *map, const float *base, int64_t rows, int64_t cols, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    PTK_REQUIRE(fn != nullptr, PTK_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable in this driver");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
    cuuint32_t box[2] = {(cuuint32_t)TG_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PTK_REQUIRE(r == CUDA_SUCCESS, PTK_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld box_rows=%d",
                (int)r, (long long)rows, (long long)cols, box_rows);
    return PTK_OK;
}

// Column tiling: N -> n_tiles CTAs of (n0 + n1) columns, each part a multiple of 16 and <= 256.
struct NTiling { int n_tiles, n0, n1; };
static NTiling plan_n(int64_t N) {
    const int64_t n16 = ceil_div(N, 16) * 16;
    NTiling t;
    if (n16 <= 256) { t.n_tiles = 1; t.n0 = (int)n16; t.n1 = 0; return t; }
    if (n16 <= TG_MAX_BN) {  // two balanced parts in one CTA (A is read and split once)
        t.n_tiles = 1;
        t.n0 = (int)(ceil_div(n16 / 2, 16) * 16);
        t.n1 = (int)(n16 - t.n0);
        return t;
    }
    t.n_tiles = (int)ceil_div(n16, 256);
    t.n0 = (int)(ceil_div(ceil_div(n16, t.n_tiles), 16) * 16);
    t.n1 = 0;
    return t;
}

bool tf32x3_eligible(const void *A, const void *D, int64_t M, int64_t K, int64_t N) {
    // TMA needs 16-byte aligned bases and row pitches; tiny reductions / outputs stay on the SIMT path
    return M >= 1 && K >= 32 && N >= 16 && (K % 4) == 0 && (((uintptr_t)A) % 16) == 0 && (((uintptr_t)D) % 16) == 0;
}

size_t tf32x3_workspace_bytes(int64_t K, int64_t N) { return 2 * sizeof(float) * (size_t)K * (size_t)N + 256; }

// D (M x N) = A (M x K) . Bsrc, where Bsrc is either (K x N) row-major [b_is_kn = 1: transposed during the
// split] or (N x K) row-major [b_is_kn = 0].  act (optional): D masked by act > 0.
int gemm_tf32x3(const float *A, const float *Bsrc, int b_is_kn, const float *act, int64_t M, int64_t K, int64_t N,
                float *D, void *workspace, size_t workspace_bytes, cudaStream_t st) {
    PTK_REQUIRE(workspace && workspace_bytes >= tf32x3_workspace_bytes(K, N), PTK_ERR_WORKSPACE,
                "gemm_tf32x3: workspace too small");
    float *b_hi = reinterpret_cast<float *>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    float *b_lo = b_hi + (size_t)K * N;
    const long long elems = (long long)K * N;
    if (b_is_kn)
        split_tf32_kernel<<<(unsigned)ceil_div(elems, 256), 256, 0, st>>>(Bsrc, (int)K, (int)N, 1, b_hi, b_lo);
    else
        split_tf32_kernel<<<(unsigned)ceil_div(elems, 256), 256, 0, st>>>(Bsrc, (int)N, (int)K, 0, b_hi, b_lo);
    PTK_CHECK_LAUNCH();

    const NTiling t = plan_n(N);
    CUtensorMap map_a, map_bhi, map_blo, map_bhi1, map_blo1;
    int rc = make_map(&map_a, A, M, K, TG_BM);
    if (rc) return rc;
    rc = make_map(&map_bhi, b_hi, N, K, t.n0);  // (N x K) K-major; box rows = width of the N-part
    if (rc) return rc;
    rc = make_map(&map_blo, b_lo, N, K, t.n0);
    if (rc) return rc;
    const int n1_box = t.n1 > 0 ? t.n1 : t.n0;
    rc = make_map(&map_bhi1, b_hi, N, K, n1_box);
    if (rc) return rc;
    rc = make_map(&map_blo1, b_lo, N, K, n1_box);
    if (rc) return rc;

    TGParams p;
    p.M = (int)M; p.N = (int)N; p.K = (int)K; p.n0_cols = t.n0; p.n1_cols = t.n1; p.D = D; p.act = act;
    const int bn = t.n0 + t.n1;
    const size_t smem = (size_t)TG_STAGES * (2 * TG_BM * TG_BK * 4 + 2 * (size_t)bn * TG_BK * 4) + 1024;
    dim3 grid((unsigned)ceil_div(M, TG_BM), (unsigned)t.n_tiles);
    if (act) {
        PTK_CHECK_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gemm_tf32x3_kernel<true><<<grid, TG_THREADS, smem, st>>>(map_a, map_bhi, map_blo, map_bhi1, map_blo1, p);
    } else {
        PTK_CHECK_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gemm_tf32x3_kernel<false><<<grid, TG_THREADS, smem, st>>>(map_a, map_bhi, map_blo, map_bhi1, map_blo1, p);
    }
    PTK_CHECK_LAUNCH();
    return PTK_OK;
}

}  // namespace ptk
