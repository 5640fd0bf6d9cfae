// Chamfer nearest-neighbour kernel template (shared by chamfer.cu and tools/chamfer_tune.cu).
// See chamfer.cu for the design notes.
#pragma once
#include "ptk_common.cuh"

namespace ptk {

constexpr int CH_TT = 2048;   // targets per shared-memory tile (3 * 2048 * 4 B = 24 KB)

__device__ __forceinline__ float min3f(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// The one and only definition of the distance arithmetic (inner loop AND index recovery).
__device__ __forceinline__ float sqdist(float qx, float qy, float qz, float tx, float ty, float tz) {
    float dx = __fsub_rn(qx, tx);
    float dy = __fsub_rn(qy, ty);
    float dz = __fsub_rn(qz, tz);
    float d = __fmul_rn(dx, dx);
    d = __fmaf_rn(dy, dy, d);
    d = __fmaf_rn(dz, dz, d);
    return d;
}

// R: queries per thread; CH_CHUNK: lazy arg-min granularity (targets); CH_THREADS: CTA size;
// MINB: resident CTAs per SM the register allocation is tuned for.
template <int R, int CH_CHUNK, int CH_THREADS, int MINB>
__global__ void __launch_bounds__(CH_THREADS, MINB)
chamfer_nn_kernel(const float *__restrict__ x, const float *__restrict__ y, int P1, int P2,
                  int split_len, int n_split, unsigned long long *__restrict__ keys_x,
                  unsigned long long *__restrict__ keys_y, int dir_only) {
    const int z = blockIdx.z;
    const int b = dir_only >= 0 ? z : (z >> 1);
    const int dir = dir_only >= 0 ? dir_only : (z & 1);
    const int NQ = dir == 0 ? P1 : P2;
    const int NT = dir == 0 ? P2 : P1;
    const float *__restrict__ Q = dir == 0 ? x + (size_t)b * P1 * 3 : y + (size_t)b * P2 * 3;
    const float *__restrict__ T = dir == 0 ? y + (size_t)b * P2 * 3 : x + (size_t)b * P1 * 3;
    unsigned long long *__restrict__ keys =
        dir == 0 ? keys_x + (size_t)b * P1 : keys_y + (size_t)b * P2;

    const int q0 = blockIdx.x * (CH_THREADS * R);
    if (q0 >= NQ) return;
    const int t_begin = blockIdx.y * split_len;
    if (t_begin >= NT) return;
    const int t_end = min(NT, t_begin + split_len);
    const int tid = threadIdx.x;

    __shared__ __align__(16) float sx[CH_TT];
    __shared__ __align__(16) float sy[CH_TT];
    __shared__ __align__(16) float sz[CH_TT];

    float qx[R], qy[R], qz[R], best[R];
    int bchunk[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        int qi = min(q0 + r * CH_THREADS + tid, NQ - 1);
        qx[r] = Q[(size_t)qi * 3 + 0];
        qy[r] = Q[(size_t)qi * 3 + 1];
        qz[r] = Q[(size_t)qi * 3 + 2];
        best[r] = __int_as_float(0x7f800000);
        bchunk[r] = t_begin / CH_CHUNK;
    }

    for (int tile = t_begin; tile < t_end; tile += CH_TT) {
        const int n = min(CH_TT, t_end - tile);
        __syncthreads();
        // global (n,3) packed floats -> shared SoA; pad to a whole chunk with +inf (never wins '<')
        const int npad = ((n + CH_CHUNK - 1) / CH_CHUNK) * CH_CHUNK;
        const float *__restrict__ src = T + (size_t)tile * 3;
        for (int e = tid; e < npad * 3; e += CH_THREADS) {
            float v = e < n * 3 ? src[e] : __int_as_float(0x7f800000);
            int p = e / 3;
            int c = e - p * 3;
            float *dst = c == 0 ? sx : (c == 1 ? sy : sz);
            dst[p] = v;
        }
        __syncthreads();
        const int nchunks = npad / CH_CHUNK;
        const int chunk0 = tile / CH_CHUNK;
        for (int c = 0; c < nchunks; ++c) {
            float m[R];
#pragma unroll
            for (int r = 0; r < R; ++r) m[r] = __int_as_float(0x7f800000);
#pragma unroll
            for (int g = 0; g < CH_CHUNK / 4; ++g) {
                const float4 tx = *reinterpret_cast<const float4 *>(&sx[c * CH_CHUNK + g * 4]);
                const float4 ty = *reinterpret_cast<const float4 *>(&sy[c * CH_CHUNK + g * 4]);
                const float4 tz = *reinterpret_cast<const float4 *>(&sz[c * CH_CHUNK + g * 4]);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    float d0 = sqdist(qx[r], qy[r], qz[r], tx.x, ty.x, tz.x);
                    float d1 = sqdist(qx[r], qy[r], qz[r], tx.y, ty.y, tz.y);
                    float d2 = sqdist(qx[r], qy[r], qz[r], tx.z, ty.z, tz.z);
                    float d3 = sqdist(qx[r], qy[r], qz[r], tx.w, ty.w, tz.w);
                    m[r] = min3f(m[r], d0, d1);
                    m[r] = min3f(m[r], d2, d3);
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                bchunk[r] = m[r] < best[r] ? chunk0 + c : bchunk[r];
                best[r] = fminf(best[r], m[r]);
            }
        }
    }

    // index recovery: first target of the recorded chunk whose distance equals the minimum
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int qi = q0 + r * CH_THREADS + tid;
        if (qi >= NQ) continue;
        const int j0 = bchunk[r] * CH_CHUNK;
        const int j1 = min(j0 + CH_CHUNK, NT);
        int arg = j0;
        bool found = false;
        for (int j = j0; j < j1; ++j) {
            float d = sqdist(qx[r], qy[r], qz[r], T[(size_t)j * 3], T[(size_t)j * 3 + 1],
                             T[(size_t)j * 3 + 2]);
            if (!found && d == best[r]) {
                arg = j;
                found = true;
            }
        }
        unsigned long long key =
            ((unsigned long long)__float_as_uint(best[r]) << 32) | (unsigned int)arg;
        if (n_split > 1)
            atomicMin(&keys[qi], key);
        else
            keys[qi] = key;
    }
}

}  // namespace ptk
