// Variant sweep for chamfer_nn_kernel (development tool, not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I include -I active-3d-vision-and-touch_b200/csrc \
//        tools/chamfer_tune.cu active-3d-vision-and-touch_b200/csrc/abi.cu -o tools/chamfer_tune
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "chamfer_kernel.cuh"

using namespace ptk;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void fill(float *p, size_t n, unsigned seed) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) {
        unsigned h = (unsigned)i * 2654435761u ^ seed;
        h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
        p[i] = (h >> 8) * (1.0f / 16777216.0f) - 0.5f;
    }
}

__global__ void checksum(const unsigned long long *k, size_t n, unsigned long long *out) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(out, k[i] * (2 * i + 1));
}

template <int R, int CHUNK, int THREADS, int MINB>
float run(const char *name, const float *x, const float *y, int B, int P, unsigned long long *kx,
          unsigned long long *ky, unsigned long long *d_sum, unsigned long long ref, int reps) {
    dim3 grid((P + THREADS * R - 1) / (THREADS * R), 1, B * 2);
    int split_len = ((P + CHUNK - 1) / CHUNK) * CHUNK;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 2; ++i)
        chamfer_nn_kernel<R, CHUNK, THREADS, MINB><<<grid, THREADS>>>(x, y, P, P, split_len, 1, kx, ky, -1);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; ++i)
        chamfer_nn_kernel<R, CHUNK, THREADS, MINB><<<grid, THREADS>>>(x, y, P, P, split_len, 1, kx, ky, -1);
    CK(cudaEventRecord(b));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    ms /= reps;
    CK(cudaMemset(d_sum, 0, 8));
    size_t n = (size_t)B * P;
    checksum<<<(unsigned)((n + 255) / 256), 256>>>(kx, n, d_sum);
    checksum<<<(unsigned)((n + 255) / 256), 256>>>(ky, n, d_sum);
    unsigned long long h;
    CK(cudaMemcpy(&h, d_sum, 8, cudaMemcpyDeviceToHost));
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, chamfer_nn_kernel<R, CHUNK, THREADS, MINB>));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, chamfer_nn_kernel<R, CHUNK, THREADS, MINB>, THREADS, 0));
    double evals = 2.0 * B * (double)P * P;
    double peak = 128.0 * 148 * 1.965e9 / 6.0;
    printf("%-28s R=%d chunk=%2d thr=%3d minb=%d regs=%3d occ=%d ctas=%5u  %8.3f ms  %6.3f Tevals/s  frac=%.3f  %s\n", name, R,
           CHUNK, THREADS, MINB, fa.numRegs, occ, grid.x * grid.z, ms, evals / ms / 1e9, evals / (ms * 1e-3) / peak,
           ref == 0 ? "(ref)" : (h == ref ? "OK" : "MISMATCH"));
    fflush(stdout);
    return ref == 0 ? (float)0 : ms;
}

int main(int argc, char **argv) {
    int B = argc > 1 ? atoi(argv[1]) : 256, P = argc > 2 ? atoi(argv[2]) : 10000, reps = argc > 3 ? atoi(argv[3]) : 5;
    size_t n = (size_t)B * P;
    float *x, *y;
    unsigned long long *kx, *ky, *d_sum;
    CK(cudaMalloc(&x, n * 12)); CK(cudaMalloc(&y, n * 12));
    CK(cudaMalloc(&kx, n * 8)); CK(cudaMalloc(&ky, n * 8)); CK(cudaMalloc(&d_sum, 8));
    fill<<<(unsigned)((n * 3 + 255) / 256), 256>>>(x, n * 3, 1u);
    fill<<<(unsigned)((n * 3 + 255) / 256), 256>>>(y, n * 3, 2u);
    CK(cudaDeviceSynchronize());
    // reference checksum from the baseline variant
    run<8, 16, 256, 2>("baseline", x, y, B, P, kx, ky, d_sum, 0, 1);
    unsigned long long ref;
    CK(cudaMemcpy(&ref, d_sum, 8, cudaMemcpyDeviceToHost));
#define V(R, C, T, M) run<R, C, T, M>(#R "," #C "," #T "," #M, x, y, B, P, kx, ky, d_sum, ref, reps)
    V(8, 16, 256, 2);
    V(8, 32, 256, 2);
    V(8, 16, 128, 4);
    V(8, 16, 128, 3);
    V(4, 16, 256, 2);
    V(4, 16, 256, 3);
    V(4, 16, 256, 4);
    V(4, 32, 256, 4);
    V(4, 16, 128, 8);
    V(6, 16, 256, 2);
    V(6, 16, 256, 3);
    V(6, 32, 256, 3);
    V(5, 16, 256, 3);
    V(3, 16, 256, 4);
    V(8, 8, 256, 2);
    V(12, 16, 128, 2);
    V(16, 16, 128, 2);
    V(8, 16, 512, 1);
    V(8, 16, 64, 8);
    return 0;
}
