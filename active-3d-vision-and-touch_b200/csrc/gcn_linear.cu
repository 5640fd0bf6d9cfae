// GCN per-vertex linear layer (FP32 GEMM) for sm_100a -- SIMT FP32 path.
//
// Replaces `torch.matmul(features, self.weight)` (pterotactyl/reconstruction/vision/model.py:352)
// and its autograd (cuBLAS sgemm in the reference):
//   fwd   : H  (M,N) = X (M,K) . W (K,N)
//   dgrad : gX (M,K) = gH (M,N) . W^T          [* (act > 0): ReLU mask of the previous layer fused]
//   wgrad : gW (K,N) = X^T (K,M) . gH (M,N)    split over M, fixed-order second-stage reduction
// Parity contract is 1e-5 relative in FP32 (BASELINE.json north_star), which rules out single-pass
// TF32/BF16 tensor-core math; this file is the exact-FP32 FFMA implementation: 128x128x16 CTA
// tiles, 8x8 register micro-tiles (2x2 blocks of 4x4 so that shared-memory reads are 128-bit and
// conflict-free), register-staged global prefetch of the next k-tile.
#include <cuda.h>
#include <stdlib.h>

#include "ptk_common.cuh"

namespace ptk {

constexpr int GL_BM = 128, GL_BN = 128, GL_BK = 16, GL_THREADS = 256, GL_PAD = 4;

// Operand addressing.  KC (k-contiguous): elem(r, k) = p[r * ld + k];  else elem(r, k) = p[k * ld + r]
template <bool KC>
__device__ __forceinline__ float ld_elem(const float *__restrict__ p, long long ld, long long r,
                                         long long k, long long R, long long K) {
    if (r >= R || k >= K) return 0.f;
    return KC ? __ldg(p + r * ld + k) : __ldg(p + k * ld + r);
}

template <bool A_KC, bool B_KC, bool MASK, bool SPLITK>
__global__ void __launch_bounds__(GL_THREADS, 2)
sgemm_kernel(const float *__restrict__ A, long long lda, const float *__restrict__ Bm, long long ldb,
             float *__restrict__ Cm, long long ldc, long long M, long long N, long long K,
             long long k_per_split, const float *__restrict__ act) {
    __shared__ __align__(16) float As[2][GL_BK][GL_BM + GL_PAD];
    __shared__ __align__(16) float Bs[2][GL_BK][GL_BN + GL_PAD];
    const int tid = threadIdx.x;
    const long long m0 = (long long)blockIdx.y * GL_BM;
    const long long n0 = (long long)blockIdx.x * GL_BN;
    const long long kb = SPLITK ? (long long)blockIdx.z * k_per_split : 0;
    const long long ke = SPLITK ? min(K, kb + k_per_split) : K;
    if (kb >= ke) return;

    float ra[8], rb[8];
    auto g2r = [&](long long k0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int idx = tid + i * GL_THREADS;
            if (A_KC) {
                const int r = idx / GL_BK, k = idx % GL_BK;
                ra[i] = ld_elem<true>(A, lda, m0 + r, k0 + k, M, ke);
            } else {
                const int k = idx / GL_BM, r = idx % GL_BM;
                ra[i] = ld_elem<false>(A, lda, m0 + r, k0 + k, M, ke);
            }
            if (B_KC) {
                const int r = idx / GL_BK, k = idx % GL_BK;
                rb[i] = ld_elem<true>(Bm, ldb, n0 + r, k0 + k, N, ke);
            } else {
                const int k = idx / GL_BN, r = idx % GL_BN;
                rb[i] = ld_elem<false>(Bm, ldb, n0 + r, k0 + k, N, ke);
            }
        }
    };
    auto r2s = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int idx = tid + i * GL_THREADS;
            if (A_KC) As[buf][idx % GL_BK][idx / GL_BK] = ra[i];
            else      As[buf][idx / GL_BM][idx % GL_BM] = ra[i];
            if (B_KC) Bs[buf][idx % GL_BK][idx / GL_BK] = rb[i];
            else      Bs[buf][idx / GL_BN][idx % GL_BN] = rb[i];
        }
    };

    const int tx = tid % 16, ty = tid / 16;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    g2r(kb);
    r2s(0);
    __syncthreads();
    int buf = 0;
    for (long long k0 = kb; k0 < ke; k0 += GL_BK) {
        const bool more = k0 + GL_BK < ke;
        if (more) g2r(k0 + GL_BK);
#pragma unroll
        for (int k = 0; k < GL_BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[buf][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[buf][k][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (more) {
            r2s(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }

#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const long long n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (n >= N) continue;
            float v = acc[i][j];
            if (MASK) v = act[m * ldc + n] > 0.f ? v : 0.f;
            Cm[(SPLITK ? (size_t)blockIdx.z * (size_t)M * (size_t)ldc : 0) + m * ldc + n] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Forward-specialised exact-FP32 kernel: H (M x N) = X (M x K) . W (K x N), all row-major, K % 4 == 0,
// N % 4 == 0, 16-byte aligned.  CTA tile BM x 160 x 16, BM = 64 by default (N = 300 -> two 160-wide column
// tiles, 6 % padding instead of the 22 % of 128-wide tiles); 8 x 10 register micro-tiles held as 8 x 5 packed
// FP32x2 pairs (40 FFMA2 per 5 shared-memory loads); 128-bit
// global loads, register-staged double buffering.  Each output is one k-sequential FMA chain starting
// from 0 -- the arithmetic of a scalar FP32 loop (see ops.py: why the training forward needs that).
constexpr int FW_BN = 160, FW_BK = 16, FW_PAD = 4;
constexpr int FW_MAXW = 16;  // mask words per row the forward can emit (K <= 512)

template <int BM>
__global__ void __launch_bounds__(BM * 2, BM <= 32 ? 6 : (BM <= 64 ? 3 : 2))
sgemm_fwd_kernel(const float *__restrict__ A, const float *__restrict__ Bm, float *__restrict__ Cm, int M, int N,
                 int K, int ld1, int nsplit, float *__restrict__ C2, int ld2, int relu2,
                 uint32_t *__restrict__ a_bits) {
    // Output: columns n < nsplit -> Cm[m * ld1 + n]; columns n >= nsplit -> C2[m * ld2 + n], optionally ReLU'd
    // (the fused GCN-layer form: the propagated slice goes to a compact scratch for the aggregation, the
    // pass-through slice straight into the layer output).  Plain GEMM: nsplit = N, ld1 = N.
    constexpr int T = BM * 2;
    pdl_launch_dependents();
    pdl_wait();
    __shared__ __align__(16) float As[2][FW_BK][BM + FW_PAD];
    __shared__ __align__(16) float Bs[2][FW_BK][FW_BN];
    // By-product for the backward: bit (k & 31) of a_bits[m * wpr + (k >> 5)] = A[m, k] > 0 -- the ReLU mask of
    // this layer's input, which the dgrad epilogue needs (every A element passes through registers here anyway).
    __shared__ uint32_t sbits[BM][FW_MAXW];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * FW_BN;
    const int tx = tid % 16, ty = tid / 16;
    const bool emit = a_bits != nullptr && blockIdx.x == 0;  // one column tile of CTAs is enough
    if (emit) {
        for (int e = tid; e < BM * FW_MAXW; e += T) sbits[e / FW_MAXW][e % FW_MAXW] = 0u;
        __syncthreads();
    }

    float4 ra[2];
    // A: global -> registers (transposed into shared memory later); B: cp.async straight into shared
    // memory (16-byte copies, zero-filled out of range) so it costs no staging registers.
    auto g2r = [&](int k0, int nbuf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int idx = tid + i * T;
            const int r = idx >> 2, kq = idx & 3;
            const int m = m0 + r, k = k0 + kq * 4;
            ra[i] = (m < M && k < K) ? __ldg(reinterpret_cast<const float4 *>(A + (size_t)m * K + k))
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < (FW_BK * 40 + T - 1) / T; ++i) {
            const int idx = tid + i * T;
            if (idx < FW_BK * 40) {
                const int kk = idx / 40, c4 = idx % 40;
                const int k = k0 + kk, n = n0 + c4 * 4;
                const bool ok = k < K && n < N;
                const float *src = ok ? Bm + (size_t)k * N + n : Bm;
                const unsigned dst = (unsigned)__cvta_generic_to_shared(&Bs[nbuf][kk][c4 * 4]);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16 : 0) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int k0_staged = 0;  // k offset of the tile held in ra
    auto r2s = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int idx = tid + i * T;
            const int r = idx >> 2, kq = idx & 3;
            if (emit) {
                // the 4 lanes kq = 0..3 of a row hold the 16 mask bits of this k-tile: OR them together with two
                // shuffles and let lane kq == 0 store the half word (no atomics)
                uint32_t h = ((ra[i].x > 0.f ? 1u : 0u) | (ra[i].y > 0.f ? 2u : 0u) | (ra[i].z > 0.f ? 4u : 0u) |
                              (ra[i].w > 0.f ? 8u : 0u)) << (kq * 4);
                h |= __shfl_xor_sync(0xffffffffu, h, 1);
                h |= __shfl_xor_sync(0xffffffffu, h, 2);
                if (kq == 0) reinterpret_cast<unsigned short *>(&sbits[r][0])[k0_staged >> 4] = (unsigned short)h;
            }
            As[buf][kq * 4 + 0][r] = ra[i].x;
            As[buf][kq * 4 + 1][r] = ra[i].y;
            As[buf][kq * 4 + 2][r] = ra[i].z;
            As[buf][kq * 4 + 3][r] = ra[i].w;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    };

    // Accumulators as packed FP32x2 pairs: FFMA2 (fma.rn.f32x2) is the same IEEE FMA per element -- the
    // k-sequential chain of every output is unchanged -- but it issues half as many instructions and reads
    // the A value as a broadcast scalar, which takes the FFMA loop off the issue/register-bank limit.
    unsigned long long acc[8][5];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) acc[i][j] = 0ull;

    g2r(0, 0);
    r2s(0);
    __syncthreads();
    int buf = 0;
    for (int k0 = 0; k0 < K; k0 += FW_BK) {
        const bool more = k0 + FW_BK < K;
        if (more) {
            g2r(k0 + FW_BK, buf ^ 1);
            k0_staged = k0 + FW_BK;
        }
#pragma unroll
        for (int k = 0; k < FW_BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][k][BM / 2 + ty * 4]);
            const ulonglong2 b0 = *reinterpret_cast<const ulonglong2 *>(&Bs[buf][k][tx * 4]);
            const ulonglong2 b1 = *reinterpret_cast<const ulonglong2 *>(&Bs[buf][k][64 + tx * 4]);
            const unsigned long long b2 = *reinterpret_cast<const unsigned long long *>(&Bs[buf][k][128 + tx * 2]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const unsigned long long bv[5] = {b0.x, b0.y, b1.x, b1.y, b2};
#pragma unroll
            for (int j = 0; j < 5; ++j)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    unsigned long long a2;
                    asm("mov.b64 %0, {%1, %1};" : "=l"(a2) : "f"(av[i]));
                    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[i][j]) : "l"(a2), "l"(bv[j]));
                }
        }
        if (more) {
            r2s(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }
    if (emit) {
        __syncthreads();
        const int wpr = (K + 31) >> 5;
        for (int e = tid; e < BM * wpr; e += T) {
            const int r = e / wpr, w = e - r * wpr;
            if (m0 + r < M) a_bits[(size_t)(m0 + r) * wpr + w] = sbits[r][w];
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : BM / 2 + ty * 4 + (i - 4));
        if (m >= M) continue;
        float *row1 = Cm + (size_t)m * ld1;
        float *row2 = C2 + (size_t)m * ld2;
        const int na = n0 + tx * 4, nb = n0 + 64 + tx * 4, nc = n0 + 128 + tx * 2;
        auto relu_pair = [&](unsigned long long v) {
            if (!relu2) return v;
            float lo = __uint_as_float((unsigned)v), hi = __uint_as_float((unsigned)(v >> 32));
            lo = fmaxf(lo, 0.f);
            hi = fmaxf(hi, 0.f);
            return ((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo);
        };
        auto put4 = [&](int n, unsigned long long a, unsigned long long b) {
            if (n + 3 >= N) return;
            if (n < nsplit)
                *reinterpret_cast<ulonglong2 *>(row1 + n) = make_ulonglong2(a, b);
            else
                *reinterpret_cast<ulonglong2 *>(row2 + n) = make_ulonglong2(relu_pair(a), relu_pair(b));
        };
        put4(na, acc[i][0], acc[i][1]);
        put4(nb, acc[i][2], acc[i][3]);
        if (nc + 1 < N) {
            if (nc < nsplit)
                *reinterpret_cast<unsigned long long *>(row1 + nc) = acc[i][4];
            else
                *reinterpret_cast<unsigned long long *>(row2 + nc) = relu_pair(acc[i][4]);
        }
    }
}

template <int BM>
static void launch_fwd(const float *X, const float *W, float *H, int64_t M, int64_t K, int64_t N, cudaStream_t st,
                       int ld1 = 0, int nsplit = -1, float *C2 = nullptr, int ld2 = 0, int relu2 = 0,
                       uint32_t *a_bits = nullptr) {
    dim3 grid((unsigned)ceil_div(N, FW_BN), (unsigned)ceil_div(M, BM));
    if (nsplit < 0) {
        nsplit = (int)N;
        ld1 = (int)N;
        C2 = H;
        ld2 = (int)N;
    }
    launch_pdl(sgemm_fwd_kernel<BM>, grid, dim3(BM * 2), 0, st, X, W, H, (int)M, (int)N, (int)K, ld1, nsplit, C2, ld2,
               relu2, K <= 32 * FW_MAXW ? a_bits : nullptr);
}

// ------------------------------------------------------------------------------------------------
// TMA-staged form of the exact forward (same arithmetic, same tile, same epilogue).  What limited the kernel above
// was not the FFMA2 stream but everything around it: per 16-wide k-tile every thread ran ~150 integer instructions
// of address / predicate arithmetic for its 2 LDG + 5 cp.async, 8 transposing STS and a CTA barrier -- ALU pipe
// 21 %, FMA pipe 75 % (profiles/r01_ncu_sgemm_fwd_bm64.txt); ALU instructions do not dual-issue with the FMA pipe
// on sm_100a.  Here one thread issues two cp.async.bulk.tensor copies per k-tile (A box 16 k x 64 rows, W box
// 160 columns x 16 k; out-of-range rows / columns / k are zero-filled by the TMA unit) into a 4-stage ring guarded by
// full / empty mbarriers; the 4 warps never meet at a CTA barrier.  A stays row-major in shared memory ([row][16 k]):
// a thread reads one float4 of 4 consecutive k per row (8 LDS.128 per 4 k-steps -- the count the transposed layout
// needed) and all 16 threads of a row group read the same address (broadcast).
constexpr int FT_BM = 64, FT_STAGES = 4, FT_THREADS = 128;
constexpr int FT_A_BYTES = FT_BM * FW_BK * 4;   // 4 KB   [64 rows][16 k]
constexpr int FT_B_BYTES = FW_BK * FW_BN * 4;   // 10 KB  [16 k][160 columns]
constexpr int FT_STAGE_BYTES = FT_A_BYTES + FT_B_BYTES;

__device__ __forceinline__ uint32_t ft_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ft_mbar_wait(uint32_t bar, uint32_t parity) {
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

// shared-space loads by 32-bit address: through C++ pointers derived from the aligned dynamic base the compiler
// loses the address space and emits generic LD.E (measured: +10 us per launch)
__device__ __forceinline__ ulonglong2 ft_lds128(uint32_t a) {
    ulonglong2 v;
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ unsigned long long ft_lds64(uint32_t a) {
    unsigned long long v;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ float2 ft_lds_f2(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ float4 ft_lds_f4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}

// One tile of RT * 8 rows x 160 columns (RT rows per thread).  RT = 8 is the main tile; smaller RT are the TAIL tiles of a
// launch: the time of this kernel is 7 us + 16 us x ceil(CTAs / 148) (tools/fwd_tail_probe.py: every SM works through its
// CTAs at one per 16 us whatever the residency), so 976 CTAs cost as much as 1036 and 912 as much as 1036 too -- 6 % and
// 14 % of the launch at the two shapes of a reconstruction step.  The rows beyond the last full round of 148 CTAs are
// therefore cut into lower tiles, chosen so that they fill one round evenly (launch_fwd_tma).  A tail tile still
// receives the 64-row TMA box (rows past its own belong to the next tile or are zero-filled) and uses its first RT * 8.
template <int RT>
__device__ __forceinline__ void ft_tile(const CUtensorMap &map_a, const CUtensorMap &map_w, float *__restrict__ Cm, int M, int N,
                                        int K, int ld1, int nsplit, float *__restrict__ C2, int ld2, int relu2,
                                        uint32_t *__restrict__ a_bits, int m0, uint32_t smem, uint32_t bar_full,
                                        uint32_t bar_empty) {
    constexpr int ROWS = RT * 8;
    const int tid = threadIdx.x, lane = tid & 31;
    const int n0 = blockIdx.x * FW_BN;
    const int tx = tid % 16, ty = tid / 16;
    // row of this thread's i-th accumulator row inside the tile (the full tile keeps its two-halves layout)
    auto row_of = [&](int i) { return RT == 8 ? (i < 4 ? ty * 4 + i : FT_BM / 2 + ty * 4 + (i - 4)) : ty * RT + i; };
    // Packed ReLU mask of A for the backward: the column tiles of a row tile share the work -- CTA x emits the k-tiles
    // with kt % gridDim.x == x (16 bits each, straight to global memory), so no CTA of a wave is slower than the others.
    const bool emit_any = a_bits != nullptr;
    const int emit_mod = (int)gridDim.x, emit_me = (int)blockIdx.x;
    const int num_kt = (K + FW_BK - 1) / FW_BK;

    auto issue = [&](int kt) {  // thread 0 only
        const int s = kt % FT_STAGES;
        const uint32_t dst = smem + (uint32_t)(s * FT_STAGE_BYTES), bar = bar_full + 8 * s;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(FT_STAGE_BYTES) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
            ::"r"(dst), "l"(&map_a), "r"(bar), "r"(kt * FW_BK), "r"(m0) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
            ::"r"(dst + FT_A_BYTES), "l"(&map_w), "r"(bar), "r"(n0), "r"(kt * FW_BK) : "memory");
    };
    if (tid == 0)
        for (int kt = 0; kt < FT_STAGES && kt < num_kt; ++kt) issue(kt);

    unsigned long long acc[RT][5];
#pragma unroll
    for (int i = 0; i < RT; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) acc[i][j] = 0ull;

    for (int kt = 0; kt < num_kt; ++kt) {
        const int s = kt % FT_STAGES;
        // refill the slot tile kt-2 lived in (one tile of slack: this thread rarely waits for the other warps)
        if (tid == 0 && kt >= 2 && kt - 2 + FT_STAGES < num_kt) {
            ft_mbar_wait(bar_empty + 8 * ((kt - 2) % FT_STAGES), ((kt - 2) / FT_STAGES) & 1);
            issue(kt - 2 + FT_STAGES);
        }
        ft_mbar_wait(bar_full + 8 * s, (kt / FT_STAGES) & 1);
        const uint32_t As = smem + (uint32_t)(s * FT_STAGE_BYTES);   // [64 rows][16 k] fp32
        const uint32_t Bs = As + FT_A_BYTES;                          // [16 k][160 columns] fp32
        if (emit_any && kt % emit_mod == emit_me) {
            // ReLU mask bits of this k-tile of A (see sgemm_fwd_kernel): thread = (row, 8-k half)
            const int r = tid >> 1, h8 = tid & 1;
            const float4 v0 = ft_lds_f4(As + 4u * (uint32_t)(r * FW_BK + h8 * 8));
            const float4 v1 = ft_lds_f4(As + 4u * (uint32_t)(r * FW_BK + h8 * 8 + 4));
            uint32_t h = ((v0.x > 0.f ? 1u : 0u) | (v0.y > 0.f ? 2u : 0u) | (v0.z > 0.f ? 4u : 0u) | (v0.w > 0.f ? 8u : 0u) |
                          (v1.x > 0.f ? 16u : 0u) | (v1.y > 0.f ? 32u : 0u) | (v1.z > 0.f ? 64u : 0u) |
                          (v1.w > 0.f ? 128u : 0u)) << (h8 * 8);
            h |= __shfl_xor_sync(0xffffffffu, h, 1);
            if (h8 == 0 && r < ROWS && m0 + r < M)
                reinterpret_cast<unsigned short *>(a_bits + (size_t)(m0 + r) * ((K + 31) >> 5))[kt] = (unsigned short)h;
        }
#pragma unroll
        for (int kq = 0; kq < FW_BK / 2; ++kq) {
            float2 a2v[RT];  // 2 k-steps of this thread's rows (float4 = 4 k-steps spilled at the 168-register cap)
#pragma unroll
            for (int i = 0; i < RT; ++i) a2v[i] = ft_lds_f2(As + 4u * (uint32_t)(row_of(i) * FW_BK + kq * 2));
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                const uint32_t brow = Bs + 4u * (uint32_t)((kq * 2 + kk) * FW_BN);
                const ulonglong2 b0 = ft_lds128(brow + 4u * (uint32_t)(tx * 4));
                const ulonglong2 b1 = ft_lds128(brow + 4u * (uint32_t)(64 + tx * 4));
                const unsigned long long b2 = ft_lds64(brow + 4u * (uint32_t)(128 + tx * 2));
                const unsigned long long bv[5] = {b0.x, b0.y, b1.x, b1.y, b2};
#pragma unroll
                for (int j = 0; j < 5; ++j)
#pragma unroll
                    for (int i = 0; i < RT; ++i) {
                        const float a = kk == 0 ? a2v[i].x : a2v[i].y;
                        unsigned long long a2;
                        asm("mov.b64 %0, {%1, %1};" : "=l"(a2) : "f"(a));
                        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[i][j]) : "l"(a2), "l"(bv[j]));
                    }
            }
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_empty + 8 * s) : "memory");
    }
    if (emit_any && (num_kt & 1) && num_kt % emit_mod == emit_me) {
        // K = 300: 19 k-tiles fill 9.5 words -- the upper half of the last word is zero (columns beyond K)
        const int wpr = (K + 31) >> 5;
        for (int r = tid; r < ROWS; r += FT_THREADS)
            if (m0 + r < M) reinterpret_cast<unsigned short *>(a_bits + (size_t)(m0 + r) * wpr)[num_kt] = 0;
    }
#pragma unroll
    for (int i = 0; i < RT; ++i) {
        const int m = m0 + row_of(i);
        if (m >= M) continue;
        float *row1 = Cm + (size_t)m * ld1;
        float *row2 = C2 + (size_t)m * ld2;
        const int na = n0 + tx * 4, nb = n0 + 64 + tx * 4, nc = n0 + 128 + tx * 2;
        auto relu_pair = [&](unsigned long long v) {
            if (!relu2) return v;
            float lo = __uint_as_float((unsigned)v), hi = __uint_as_float((unsigned)(v >> 32));
            lo = fmaxf(lo, 0.f);
            hi = fmaxf(hi, 0.f);
            return ((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo);
        };
        auto put4 = [&](int n, unsigned long long a, unsigned long long b) {
            if (n + 3 >= N) return;
            if (n < nsplit)
                *reinterpret_cast<ulonglong2 *>(row1 + n) = make_ulonglong2(a, b);
            else
                *reinterpret_cast<ulonglong2 *>(row2 + n) = make_ulonglong2(relu_pair(a), relu_pair(b));
        };
        put4(na, acc[i][0], acc[i][1]);
        put4(nb, acc[i][2], acc[i][3]);
        if (nc + 1 < N) {
            if (nc < nsplit)
                *reinterpret_cast<unsigned long long *>(row1 + nc) = acc[i][4];
            else
                *reinterpret_cast<unsigned long long *>(row2 + nc) = relu_pair(acc[i][4]);
        }
    }
}

// grid (column tiles, n_main + n_tail): row tiles [0, n_main) are 64 rows high, the n_tail tiles behind them RT_TAIL * 8.
template <int RT_TAIL>
__global__ void __launch_bounds__(FT_THREADS, 3)
sgemm_fwd_tma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                     float *__restrict__ Cm, int M, int N, int K, int ld1, int nsplit, float *__restrict__ C2, int ld2,
                     int relu2, uint32_t *__restrict__ a_bits, int n_main) {
    extern __shared__ __align__(128) uint8_t ft_smem_raw[];
    const uint32_t smem = (ft_smem_u32(ft_smem_raw) + 127u) & ~127u;  // shared-space address of stage 0
    __shared__ __align__(8) uint64_t bars[2 * FT_STAGES];
    const uint32_t bar_full = ft_smem_u32(&bars[0]), bar_empty = ft_smem_u32(&bars[FT_STAGES]);
    pdl_launch_dependents();
    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
        for (int s = 0; s < FT_STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_full + 8 * s), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_empty + 8 * s), "r"(FT_THREADS / 32));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_wait();  // A is the predecessor's output
    const int y = (int)blockIdx.y;
    if (RT_TAIL == 8 || y < n_main)
        ft_tile<8>(map_a, map_w, Cm, M, N, K, ld1, nsplit, C2, ld2, relu2, a_bits, y * FT_BM, smem, bar_full, bar_empty);
    else
        ft_tile<RT_TAIL>(map_a, map_w, Cm, M, N, K, ld1, nsplit, C2, ld2, relu2, a_bits,
                         n_main * FT_BM + (y - n_main) * RT_TAIL * 8, smem, bar_full, bar_empty);
}

int make_tensor_map_2d(CUtensorMap *map, const float *base, int64_t rows, int64_t cols, int box_cols, int box_rows,
                       bool swizzle);  // gemm_tf32x3.cu

static int launch_fwd_tma(const float *X, const float *W, float *H, int64_t M, int64_t K, int64_t N, cudaStream_t st,
                          int ld1 = 0, int nsplit = -1, float *C2 = nullptr, int ld2 = 0, int relu2 = 0,
                          uint32_t *a_bits = nullptr) {
    CUtensorMap map_a, map_w;
    int rc = make_tensor_map_2d(&map_a, X, M, K, FW_BK, FT_BM, false);
    if (rc) return rc;
    rc = make_tensor_map_2d(&map_w, W, K, N, FW_BN, FW_BK, false);
    if (rc) return rc;
    if (nsplit < 0) {
        nsplit = (int)N;
        ld1 = (int)N;
        C2 = H;
        ld2 = (int)N;
    }
    const size_t smem = (size_t)FT_STAGES * FT_STAGE_BYTES + 128;
    int dev = 0;
    PTK_CHECK_CUDA(cudaGetDevice(&dev));
    // Tail plan: rows behind the last FULL round of sm_count CTAs go into lower tiles (RT * 8 rows) that fill one more
    // round as evenly as they can; cost model = rounds x tile height, small tiles charged 3 % per missing row group
    // (fewer FFMA2 per shared-memory load).  PTK_FWD_TAIL=0 keeps 64-row tiles everywhere (development).
    const int64_t cx = ceil_div(N, FW_BN), tiles64 = ceil_div(M, (int64_t)FT_BM), Q = sm_count();
    int64_t n_main = tiles64;
    int rt_tail = 8;
    static const bool tail_on = !(getenv("PTK_FWD_TAIL") && atoi(getenv("PTK_FWD_TAIL")) == 0);
    if (tail_on && (cx * tiles64) % Q != 0) {
        n_main = (cx * tiles64 / Q) * Q / cx;  // row tiles of the full rounds
        const int64_t rem = M - n_main * FT_BM;
        double best = 1e30;
        for (int rt = 8; rt >= 1; --rt) {
            const double cost = (double)ceil_div(cx * ceil_div(rem, (int64_t)rt * 8), Q) * rt * (1.0 + 0.03 * (8 - rt));
            if (cost < best - 1e-9) {
                best = cost;
                rt_tail = rt;
            }
        }
        if (rt_tail == 8) n_main = tiles64;
    }
    const int64_t n_tail = rt_tail == 8 ? 0 : ceil_div(M - n_main * FT_BM, (int64_t)rt_tail * 8);
    dim3 grid((unsigned)cx, (unsigned)(n_main + n_tail));
    uint32_t *bits = K <= 32 * FW_MAXW ? a_bits : nullptr;
#define PTK_FT_LAUNCH(RT)                                                                                                  \
    do {                                                                                                                   \
        static unsigned long long optin = 0ull; /* cudaFuncAttributeMaxDynamicSharedMemorySize is per device */             \
        if (dev >= 64 || !((optin >> dev) & 1ull)) {                                                                       \
            PTK_CHECK_CUDA(cudaFuncSetAttribute(sgemm_fwd_tma_kernel<RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            if (dev < 64) optin |= 1ull << dev;                                                                            \
        }                                                                                                                  \
        launch_pdl(sgemm_fwd_tma_kernel<RT>, grid, dim3(FT_THREADS), smem, st, map_a, map_w, H, (int)M, (int)N, (int)K, ld1,   \
                   nsplit, C2, ld2, relu2, bits, (int)n_main);                                                             \
    } while (0)
    switch (rt_tail) {
        case 1: PTK_FT_LAUNCH(1); break;
        case 2: PTK_FT_LAUNCH(2); break;
        case 3: PTK_FT_LAUNCH(3); break;
        case 4: PTK_FT_LAUNCH(4); break;
        case 5: PTK_FT_LAUNCH(5); break;
        case 6: PTK_FT_LAUNCH(6); break;
        case 7: PTK_FT_LAUNCH(7); break;
        default: PTK_FT_LAUNCH(8); break;
    }
#undef PTK_FT_LAUNCH
    PTK_CHECK_LAUNCH();
    return PTK_OK;
}

// ------------------------------------------------------------------------------------------------
// Skinny layers (N <= 4: the 300 -> 3 output layer of every GCN, vision/model.py:296-301).  As GEMMs they are
// matrix-vector shaped and HBM/L2-bound; the tiled kernels above spend 80 us on them, these take ~10 us.
//   forward : one warp per row, 128-bit loads of X, W (K x N) in shared memory, warp reduction
//   wgrad   : a CTA sums a slab of rows for every k (thread = k), partials reduced by splitk_reduce_kernel
// (The output layer has no activation behind it, so the k-sequential rounding of the exact forward kernel is
// not needed here.)
constexpr int SK_MAXN = 4;

__global__ void __launch_bounds__(256)
skinny_fwd_kernel(const float *__restrict__ X, const float *__restrict__ W, float *__restrict__ H, long long M, int K,
                  int N) {
    extern __shared__ __align__(16) float sk_w[];  // [K][4]
    for (int e = threadIdx.x; e < K * SK_MAXN; e += 256) {
        const int k = e >> 2, n = e & 3;
        sk_w[e] = n < N ? W[(size_t)k * N + n] : 0.f;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= M) return;
    const float4 *x4 = reinterpret_cast<const float4 *>(X + (size_t)row * K);
    float acc[SK_MAXN] = {0.f, 0.f, 0.f, 0.f};
    for (int q = lane; q < K / 4; q += 32) {
        const float4 x = __ldg(x4 + q);
        const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 w = *reinterpret_cast<const float4 *>(&sk_w[(q * 4 + j) * SK_MAXN]);
            acc[0] = fmaf(xv[j], w.x, acc[0]); acc[1] = fmaf(xv[j], w.y, acc[1]);
            acc[2] = fmaf(xv[j], w.z, acc[2]); acc[3] = fmaf(xv[j], w.w, acc[3]);
        }
    }
#pragma unroll
    for (int n = 0; n < SK_MAXN; ++n) acc[n] = warp_sum(acc[n]);
    if (lane < N) H[(size_t)row * N + lane] = lane == 0 ? acc[0] : (lane == 1 ? acc[1] : (lane == 2 ? acc[2] : acc[3]));
}

// part[blockIdx.x][k][n] = sum over the CTA's rows of X[m,k] * gH[m,n]
__global__ void __launch_bounds__(1024)
skinny_wgrad_kernel(const float *__restrict__ X, const float *__restrict__ gH, float *__restrict__ part, long long M,
                    int K, int N, long long rows_per_cta) {
    const long long r0 = (long long)blockIdx.x * rows_per_cta;
    const long long r1 = r0 + rows_per_cta < M ? r0 + rows_per_cta : M;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        float acc[SK_MAXN] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
        for (long long m = r0; m < r1; ++m) {
            const float x = X[(size_t)m * K + k];
#pragma unroll
            for (int n = 0; n < SK_MAXN; ++n)
                if (n < N) acc[n] = fmaf(x, __ldg(gH + (size_t)m * N + n), acc[n]);
        }
        for (int n = 0; n < N; ++n) part[((size_t)blockIdx.x * K + k) * N + n] = acc[n];
    }
}

// second stage of the split wgrad: out[e] = sum_s part[s, e] in ascending s (deterministic)
__global__ void splitk_reduce_kernel(const float *__restrict__ part, int nsplit, long long elems,
                                     float *__restrict__ out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= elems) return;
    float acc = 0.f;
    for (int s = 0; s < nsplit; ++s) acc += part[(size_t)s * elems + e];
    out[e] = acc;
}

// tensor-core path (gemm_tf32x3.cu)
bool tf32x3_eligible(const void *A, const void *D, int64_t M, int64_t K, int64_t N);
size_t tf32x3_workspace_bytes(int64_t M, int64_t K, int64_t N);
int gemm_tf32x3(const float *A, const float *Bsrc, int b_is_kn, const float *act, int64_t M, int64_t K, int64_t N,
                float *D, void *workspace, size_t workspace_bytes, cudaStream_t st, const uint32_t *act_bits = nullptr);

bool wgrad_tf32x3_eligible(const void *X, const void *gH, int64_t M, int64_t Kin, int64_t Nout);
size_t wgrad_tf32x3_workspace_bytes(int64_t M, int64_t Kin, int64_t Nout);
int wgrad_tf32x3(const float *X, const float *gH, int64_t M, int64_t Kin, int64_t Nout, void *workspace,
                 size_t workspace_bytes, float **part_out, int *n_splits, cudaStream_t st);

}  // namespace ptk

using namespace ptk;

static int check_gemm(const void *a, const void *b, const void *c, int64_t M, int64_t K, int64_t N) {
    PTK_REQUIRE(a && b && c, PTK_ERR_SHAPE, "gcn_linear: null pointer");
    PTK_REQUIRE(M > 0 && K > 0 && N > 0, PTK_ERR_SHAPE, "gcn_linear: bad sizes (M=%lld, K=%lld, N=%lld)",
                (long long)M, (long long)K, (long long)N);
    PTK_REQUIRE(ceil_div(M, GL_BM) <= 65535, PTK_ERR_SHAPE, "gcn_linear: M too large for one launch");
    return PTK_OK;
}

extern "C" size_t ptk_gcn_linear_workspace_bytes(int64_t M, int64_t K, int64_t N) {
    if (M <= 0 || K <= 0 || N <= 0) return 0;
    return tf32x3_workspace_bytes(M, K, N);
}

extern "C" int ptk_gcn_linear_fwd(const float *X, const float *W, int64_t M, int64_t K, int64_t N,
                                  float *H, int algo, void *workspace, size_t workspace_bytes,
                                  ptk_stream_t stream) {
    PTK_NVTX("ptk_gcn_linear_fwd");
    int rc = check_gemm(X, W, H, M, K, N);
    if (rc) return rc;
    PTK_REQUIRE(algo >= 0 && algo <= 2, PTK_ERR_SHAPE, "gcn_linear_fwd: algo must be 0, 1 or 2");
    const int g_fwd_mode = algo;
    if (g_fwd_mode != 1 && tf32x3_eligible(X, H, M, K, N))
        return gemm_tf32x3(X, W, /*b_is_kn=*/1, nullptr, M, K, N, H, workspace, workspace_bytes, as_stream(stream));
    PTK_REQUIRE(g_fwd_mode != 2, PTK_ERR_SHAPE, "gcn_linear_fwd: shape not eligible for the tensor-core path");
    if ((K % 4) == 0 && (N % 4) == 0 && N >= 64 && (((uintptr_t)X | (uintptr_t)W | (uintptr_t)H) % 16) == 0 &&
        M < (1LL << 31) && ceil_div(M, 32) <= 65535) {
        // BM = 64 (128 threads, 3 CTAs/SM, 168 registers: the 8 x 10 micro-tile fits without spills) measured
        // fastest at every M that occurs here: 120 us vs 145 us (BM 112/128, spilling at the 128-register cap),
        // 136 us (BM 32) and cuBLAS FP32 127 us at M=31184, K=N=300.
        int best = 64;
        if (PTK_TUNING_ENV("PTK_FWD_BM") > 0) best = PTK_TUNING_ENV("PTK_FWD_BM");  // tools/gemm_check.py sweeps
        if (best == 64 && !PTK_TUNING_ENV("PTK_FWD_NO_TMA")) return launch_fwd_tma(X, W, H, M, K, N, as_stream(stream));
        if (best == 32) launch_fwd<32>(X, W, H, M, K, N, as_stream(stream));
        else if (best == 64) launch_fwd<64>(X, W, H, M, K, N, as_stream(stream));
        else if (best == 128) launch_fwd<128>(X, W, H, M, K, N, as_stream(stream));
        else if (best == 112) launch_fwd<112>(X, W, H, M, K, N, as_stream(stream));
        else launch_fwd<96>(X, W, H, M, K, N, as_stream(stream));
        PTK_CHECK_LAUNCH();
        return PTK_OK;
    }
    if (N <= SK_MAXN && (K % 4) == 0 && ((uintptr_t)X % 16) == 0 && (size_t)K * SK_MAXN * 4 <= 48 * 1024) {
        skinny_fwd_kernel<<<(unsigned)ceil_div(M, 8), 256, (size_t)K * SK_MAXN * sizeof(float), as_stream(stream)>>>(
            X, W, H, (long long)M, (int)K, (int)N);
        PTK_CHECK_LAUNCH();
        return PTK_OK;
    }
    dim3 grid((unsigned)ceil_div(N, GL_BN), (unsigned)ceil_div(M, GL_BM));
    sgemm_kernel<true, false, false, false><<<grid, GL_THREADS, 0, as_stream(stream)>>>(
        X, K, W, N, H, N, M, N, K, 0, nullptr);
    PTK_CHECK_LAUNCH();
    return PTK_OK;
}

extern "C" int ptk_gcn_linear_fwd_split(const float *X, const float *W, int64_t M, int64_t K, int64_t N,
                                        int64_t n_split, float *head, float *out, int relu, uint32_t *x_bits,
                                        ptk_stream_t stream) {
    PTK_NVTX("ptk_gcn_linear_fwd_split");
    int rc = check_gemm(X, W, out, M, K, N);
    if (rc) return rc;
    PTK_REQUIRE(head && n_split > 0 && n_split < N && (n_split % 4) == 0, PTK_ERR_SHAPE,
                "gcn_linear_fwd_split: n_split must be a multiple of 4 inside (0, N)");
    PTK_REQUIRE((K % 4) == 0 && (N % 4) == 0 && N >= 64 && M < (1LL << 31) && ceil_div(M, 64) <= 65535, PTK_ERR_SHAPE,
                "gcn_linear_fwd_split: needs K %% 4 == 0, N %% 4 == 0, N >= 64");
    PTK_REQUIRE((((uintptr_t)X | (uintptr_t)W | (uintptr_t)head | (uintptr_t)out) % 16) == 0, PTK_ERR_ALIGN,
                "gcn_linear_fwd_split: pointers must be 16-byte aligned");
    PTK_REQUIRE(!x_bits || K <= 32 * FW_MAXW, PTK_ERR_SHAPE, "gcn_linear_fwd_split: x_bits needs K <= %d", 32 * FW_MAXW);
    if (!PTK_TUNING_ENV("PTK_FWD_NO_TMA"))
        return launch_fwd_tma(X, W, head, M, K, N, as_stream(stream), (int)n_split, (int)n_split, out, (int)N, relu, x_bits);
    launch_fwd<64>(X, W, head, M, K, N, as_stream(stream), (int)n_split, (int)n_split, out, (int)N, relu, x_bits);
    PTK_CHECK_LAUNCH();
    return PTK_OK;
}

extern "C" int ptk_gcn_linear_dgrad(const float *gH, const float *W, const float *act, const uint32_t *act_bits,
                                    int64_t M, int64_t K, int64_t N, float *gX, int algo, void *workspace,
                                    size_t workspace_bytes, ptk_stream_t stream) {
    PTK_NVTX("ptk_gcn_linear_dgrad");
    int rc = check_gemm(gH, W, gX, M, K, N);
    if (rc) return rc;
    PTK_REQUIRE(algo >= 0 && algo <= 2, PTK_ERR_SHAPE, "gcn_linear_dgrad: algo must be 0, 1 or 2");
    const int g_dgrad_mode = algo;
    // gX (M x K) = gH (M x N) . W^T : reduction over N; W as stored (K x N) is the "N' x K'" operand
    if (g_dgrad_mode != 1 && tf32x3_eligible(gH, gX, M, N, K) && (!act || ((uintptr_t)act % 16) == 0))
        return gemm_tf32x3(gH, W, /*b_is_kn=*/0, act, M, N, K, gX, workspace, workspace_bytes, as_stream(stream),
                           act ? act_bits : nullptr);
    PTK_REQUIRE(g_dgrad_mode != 2, PTK_ERR_SHAPE, "gcn_linear_dgrad: shape not eligible for the tensor-core path");
    // gX (M x K) = gH (M x N) . W^T : GEMM with m=M, n=K, k=N; B[k=n_out][n=k_in] = W[k_in*N + n_out]
    dim3 grid((unsigned)ceil_div(K, GL_BN), (unsigned)ceil_div(M, GL_BM));
    if (act)
        sgemm_kernel<true, true, true, false><<<grid, GL_THREADS, 0, as_stream(stream)>>>(
            gH, N, W, N, gX, K, M, K, N, 0, act);
    else
        sgemm_kernel<true, true, false, false><<<grid, GL_THREADS, 0, as_stream(stream)>>>(
            gH, N, W, N, gX, K, M, K, N, 0, nullptr);
    PTK_CHECK_LAUNCH();
    return PTK_OK;
}

static int skinny_wgrad_ctas(int64_t M) {
    const int64_t want = 3LL * sm_count();  // three 320-thread CTAs per SM keep enough loads in flight
    const int64_t n = M / 16 < want ? M / 16 : want;
    return (int)(n > 0 ? n : 1);
}

static int wgrad_splits(int64_t M, int64_t K, int64_t N) {
    if (N <= SK_MAXN) return skinny_wgrad_ctas(M);
    const int64_t tiles = ceil_div(K, GL_BM) * ceil_div(N, GL_BN);
    int64_t want = ceil_div(2LL * sm_count() * 2, tiles);
    int64_t max_s = ceil_div(M, 4 * GL_BK);
    if (want > max_s) want = max_s;
    if (want < 1) want = 1;
    return (int)want;
}

extern "C" size_t ptk_gcn_linear_wgrad_workspace_bytes(int64_t M, int64_t K, int64_t N) {
    if (M <= 0 || K <= 0 || N <= 0) return 0;
    const size_t a = sizeof(float) * (size_t)wgrad_splits(M, K, N) * (size_t)K * (size_t)N;
    const size_t b = wgrad_tf32x3_workspace_bytes(M, K, N);
    return a > b ? a : b;
}

extern "C" int ptk_gcn_linear_wgrad(const float *X, const float *gH, int64_t M, int64_t K, int64_t N,
                                    float *gW, int algo, void *workspace, size_t workspace_bytes,
                                    ptk_stream_t stream) {
    PTK_NVTX("ptk_gcn_linear_wgrad");
    int rc = check_gemm(X, gH, gW, M, K, N);
    if (rc) return rc;
    PTK_REQUIRE(algo >= 0 && algo <= 2, PTK_ERR_SHAPE, "gcn_linear_wgrad: algo must be 0, 1 or 2");
    PTK_REQUIRE(workspace && workspace_bytes >= ptk_gcn_linear_wgrad_workspace_bytes(M, K, N),
                PTK_ERR_WORKSPACE, "gcn_linear_wgrad: workspace too small");
    if (algo != 1 && wgrad_tf32x3_eligible(X, gH, M, K, N)) {
        float *part = nullptr;
        int ns = 0;
        rc = wgrad_tf32x3(X, gH, M, K, N, workspace, workspace_bytes, &part, &ns, as_stream(stream));
        if (rc) return rc;
        const long long elems = (long long)K * N;
        splitk_reduce_kernel<<<(unsigned)ceil_div(elems, 256), 256, 0, as_stream(stream)>>>(part, ns, elems, gW);
        PTK_CHECK_LAUNCH();
        return PTK_OK;
    }
    PTK_REQUIRE(algo != 2, PTK_ERR_SHAPE, "gcn_linear_wgrad: shape not eligible for the tensor-core path");
    if (N <= SK_MAXN) {
        const int ctas = skinny_wgrad_ctas(M);
        const long long rows_per_cta = (long long)ceil_div(M, ctas);
        float *part = reinterpret_cast<float *>(workspace);
        cudaStream_t st = as_stream(stream);
        const int threads = (int)(K >= 1024 ? 1024 : ceil_div(K, 32) * 32);
        skinny_wgrad_kernel<<<ctas, threads, 0, st>>>(X, gH, part, (long long)M, (int)K, (int)N, rows_per_cta);
        PTK_CHECK_LAUNCH();
        const long long elems = (long long)K * N;
        splitk_reduce_kernel<<<(unsigned)ceil_div(elems, 256), 256, 0, st>>>(part, ctas, elems, gW);
        PTK_CHECK_LAUNCH();
        return PTK_OK;
    }
    // gW (K x N) = X^T . gH : GEMM with m=K, n=N, k=M; A[m=k_in][k=row] = X[row*K + k_in]
    const int ns = wgrad_splits(M, K, N);
    const int64_t kper = ceil_div(ceil_div(M, ns), GL_BK) * GL_BK;
    const int ns_eff = (int)ceil_div(M, kper);
    float *part = reinterpret_cast<float *>(workspace);
    dim3 grid((unsigned)ceil_div(N, GL_BN), (unsigned)ceil_div(K, GL_BM), (unsigned)ns_eff);
    cudaStream_t st = as_stream(stream);
    sgemm_kernel<false, false, false, true><<<grid, GL_THREADS, 0, st>>>(X, K, gH, N, part, N, K, N, M,
                                                                         kper, nullptr);
    PTK_CHECK_LAUNCH();
    const long long elems = (long long)K * N;
    splitk_reduce_kernel<<<(unsigned)ceil_div(elems, 256), 256, 0, st>>>(part, ns_eff, elems, gW);
    PTK_CHECK_LAUNCH();
    return PTK_OK;
}
