// Dev probe: issue rate of packed FP32x2 instructions on sm_100a (cycles per warp instruction per SMSP).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/fp32x2_probe.cu -o tools/fp32x2_probe
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float min3(float a, float b, float c) { float r; asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

constexpr int NACC = 16, ITERS = 2048;

template <int MODE>
__global__ void probe(float *out, long long *cyc, float s0, float s1) {
    u64 acc[NACC];
    float facc[NACC];
    float q[8];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { acc[i] = pack2(threadIdx.x * 1e-3f + i, i * 0.5f); facc[i] = threadIdx.x * 1e-3f + i; }
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = s0 + i * 1e-3f + threadIdx.x * 1e-6f;
    u64 b2 = pack2(s0, s1), c2 = pack2(s1, s0);
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (MODE == 0) facc[i] = fma1(facc[i], s0, s1);                 // FFMA scalar (2 ops: counts NACC)
            if (MODE == 1) acc[i] = fma2(acc[i], b2, c2);                   // FFMA2 acc*b+c (b,c shared)
            if (MODE == 2) acc[i] = fma2(b2, c2, acc[i]);                   // FFMA2 b*c+acc
            if (MODE == 3) acc[i] = fma2(pack2(q[i & 7], q[i & 7]), b2, acc[i]);  // broadcast scalar * shared pair + acc
            if (MODE == 4) acc[i] = add2(acc[i], b2);                       // FADD2
            if (MODE == 5) acc[i] = mul2(acc[i], b2);                       // FMUL2
            if (MODE == 6) { acc[i] = fma2(pack2(q[i & 7], q[i & 7]), b2, acc[i]); facc[i] = fma1(q[i & 7], s0, facc[i]); }  // FFMA2 + FFMA
            if (MODE == 7) { acc[i] = fma2(pack2(q[i & 7], q[i & 7]), b2, acc[i]); facc[i] = min3(facc[i], q[i & 7], s1); }  // FFMA2 + FMNMX3
            if (MODE == 8) facc[i] = fma1(q[i & 7], s0, facc[i]);           // FFMA q*s+acc
            if (MODE == 9) acc[i] = fma2(acc[i], acc[(i + 1) % NACC], acc[(i + 2) % NACC]);  // 3 distinct pairs
            if (MODE == 10) { facc[i] = fma1(q[i & 7], s0, facc[i]); float t = min3(__uint_as_float((unsigned)(acc[i] >> 32)), facc[i], s1); acc[i] = pack2(t, t); } // FFMA + FMNMX3
        }
    }
    long long t1 = clock64();
    float r = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) { r += facc[i] + __uint_as_float((unsigned)acc[i]) + __uint_as_float((unsigned)(acc[i] >> 32)); }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int instr_per_iter, int warps_per_smsp) {
    float *out; long long *cyc;
    int threads = 128 * warps_per_smsp, blocks = 148;
    cudaMalloc(&out, sizeof(float) * threads * blocks); cudaMalloc(&cyc, 8 * blocks);
    probe<MODE><<<blocks, threads>>>(out, cyc, 1.0001f, 0.5f);
    probe<MODE><<<blocks, threads>>>(out, cyc, 1.0001f, 0.5f);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, 8 * blocks, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
    double per = avg / ((double)ITERS * instr_per_iter * warps_per_smsp);
    printf("%-44s warps/SMSP=%d  %.3f cycles per warp-instr per SMSP  (err %s)\n", name, warps_per_smsp, per, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {1, 2, 4}) {
        run<0>("FFMA acc*s+s", NACC, w);
        run<8>("FFMA q*s+acc", NACC, w);
        run<1>("FFMA2 acc*b+c (b,c shared pairs)", NACC, w);
        run<2>("FFMA2 b*c+acc", NACC, w);
        run<3>("FFMA2 q.F32*b+acc", NACC, w);
        run<9>("FFMA2 3 distinct pairs", NACC, w);
        run<4>("FADD2", NACC, w);
        run<5>("FMUL2", NACC, w);
        run<6>("FFMA2 + FFMA (per instr)", 2 * NACC, w);
        run<7>("FFMA2 + FMNMX3 (per instr)", 2 * NACC, w);
        run<10>("FFMA + FMNMX3 (per instr)", 2 * NACC, w);
    }
    return 0;
}
