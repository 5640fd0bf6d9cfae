// Max over the vertices of per-vertex features: (B, Nv, C) -> (B, C), with the arg-max for the backward.
//
// Replaces `features.max(dim=1)[0]` after the GCN encoder of the autoencoder
// (pterotactyl/reconstruction/autoencoder/model.py:91) and `torch.max(x, dim=1)[0]` of the DDQN
// Graph_Model (pterotactyl/policies/DDQN/model.py:128).  HBM-bound: every input byte is read once.
// CTA = (32-channel slab, batch element); 8 row lanes x 32 channel lanes, 128-byte coalesced rows;
// fixed-order shared-memory reduction over the row lanes => deterministic, lowest vertex wins ties,
// NaN propagates like torch.max.
#include "ptk_common.cuh"

namespace ptk {

__device__ __forceinline__ bool better(float v, int i, float bv, int bi) {
    // v replaces bv when it is larger, or NaN and bv is not, or equal with a lower vertex id
    const bool vn = v != v, bn = bv != bv;
    if (vn || bn) return vn && (!bn || i < bi);
    return v > bv || (v == bv && i < bi);
}

__global__ void __launch_bounds__(256)
vertex_maxpool_kernel(const float *__restrict__ in, int Nv, int C, float *__restrict__ out,
                      int32_t *__restrict__ arg) {
    __shared__ float sv[8][33];
    __shared__ int si[8][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cx;
    const int b = blockIdx.y;
    const float *base = in + (size_t)b * Nv * C;
    float bv = __int_as_float(0xff800000);  // -inf
    int bi = 0x7fffffff;
    if (c < C) {
        for (int i = ry; i < Nv; i += 8) {
            const float v = base[(size_t)i * C + c];
            if (better(v, i, bv, bi)) {
                bv = v;
                bi = i;
            }
        }
    }
    sv[ry][cx] = bv;
    si[ry][cx] = bi;
    __syncthreads();
    if (ry == 0 && c < C) {
#pragma unroll
        for (int y = 1; y < 8; ++y)
            if (better(sv[y][cx], si[y][cx], bv, bi)) {
                bv = sv[y][cx];
                bi = si[y][cx];
            }
        out[(size_t)b * C + c] = bv;
        if (arg) arg[(size_t)b * C + c] = bi;
    }
}

// grad_in (zero-filled by the caller's memset below) [b, arg[b,c], c] = grad_out[b, c]
__global__ void vertex_maxpool_bwd_kernel(const float *__restrict__ gout, const int32_t *__restrict__ arg, int Nv,
                                          int C, long long n, float *__restrict__ gin) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const long long b = e / C;
    const int c = (int)(e - b * C);
    gin[((size_t)b * Nv + arg[e]) * C + c] = gout[e];
}

}  // namespace ptk

using namespace ptk;

extern "C" int ptk_vertex_maxpool_fwd(const float *in, int64_t B, int64_t Nv, int64_t C, float *out, int32_t *arg,
                                      ptk_stream_t stream) {
    PTK_REQUIRE(in && out, PTK_ERR_SHAPE, "vertex_maxpool_fwd: null pointer");
    PTK_REQUIRE(B > 0 && Nv > 0 && C > 0 && B <= 65535 && Nv < (1LL << 31), PTK_ERR_SHAPE,
                "vertex_maxpool_fwd: bad sizes (B=%lld, Nv=%lld, C=%lld)", (long long)B, (long long)Nv, (long long)C);
    dim3 grid((unsigned)ceil_div(C, 32), (unsigned)B);
    vertex_maxpool_kernel<<<grid, 256, 0, as_stream(stream)>>>(in, (int)Nv, (int)C, out, arg);
    PTK_CHECK_LAUNCH();
    return PTK_OK;
}

extern "C" int ptk_vertex_maxpool_bwd(const float *grad_out, const int32_t *arg, int64_t B, int64_t Nv, int64_t C,
                                      float *grad_in, ptk_stream_t stream) {
    PTK_REQUIRE(grad_out && arg && grad_in, PTK_ERR_SHAPE, "vertex_maxpool_bwd: null pointer");
    PTK_REQUIRE(B > 0 && Nv > 0 && C > 0, PTK_ERR_SHAPE, "vertex_maxpool_bwd: bad sizes");
    cudaStream_t st = as_stream(stream);
    PTK_CHECK_CUDA(cudaMemsetAsync(grad_in, 0, sizeof(float) * (size_t)B * Nv * C, st));
    const long long n = (long long)B * C;
    vertex_maxpool_bwd_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(grad_out, arg, (int)Nv, (int)C, n, grad_in);
    PTK_CHECK_LAUNCH();
    return PTK_OK;
}
