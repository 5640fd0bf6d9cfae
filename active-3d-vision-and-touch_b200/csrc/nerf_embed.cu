// NeRF positional embedding of the vertex positions, fused: (M,3) -> (M,63) in one launch.
//
// Replaces Positional_Encoder.nerf_embedding + the concatenation with the raw positions
// (pterotactyl/reconstruction/vision/model.py:381-391, 396-397): 20 sin/cos launches, 20 scalar multiplies and two
// torch.cat per call, three calls per forward (model.py:229, 262, 274).  Column layout of a row, as the reference's
// cat order gives it:  [6 i + c] = sin(s_i * p_c),  [6 i + 3 + c] = cos(s_i * p_c)  (i = 0..9, c = x,y,z),
// [60 + c] = p_c, with s_0 = fl32(pi) and s_i = fl32(pi * 2 * i) -- a Python double rounded once when it meets the
// fp32 tensor (model.py:385-389; note 2*i, not 2^i).  sinf / cosf are the accurate libdevice functions torch's
// CUDA sin / cos call, so the forward agrees with the reference on a GPU bit for bit and with its CPU run to 1 ulp.
// Backward: g_p = g[60 + c] + sum_i s_i * (cos(s_i p) * g_sin - sin(s_i p) * g_cos), one thread per coordinate.
#include "ptk_common.cuh"

namespace ptk {

constexpr int NE_FREQ = 10, NE_W = 6 * NE_FREQ + 3;

__device__ __forceinline__ float nerf_scale(int i) {
    // fl32 of the Python doubles np.pi (i = 0) and np.pi * 2 * i
    return i == 0 ? (float)3.141592653589793 : (float)((3.141592653589793 * 2.0) * (double)i);
}

__global__ void __launch_bounds__(256)
nerf_embed_fwd_kernel(const float *__restrict__ pos, long long M, float *__restrict__ out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= M * NE_W) return;
    const long long row = e / NE_W;
    const int c = (int)(e - row * NE_W);
    if (c >= 6 * NE_FREQ) {
        out[e] = pos[row * 3 + (c - 6 * NE_FREQ)];
        return;
    }
    const int i = c / 6, r = c - 6 * i;
    const float arg = __fmul_rn(nerf_scale(i), pos[row * 3 + (r < 3 ? r : r - 3)]);
    out[e] = r < 3 ? sinf(arg) : cosf(arg);
}

__global__ void __launch_bounds__(256)
nerf_embed_bwd_kernel(const float *__restrict__ pos, const float *__restrict__ gout, long long M,
                      float *__restrict__ gpos) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= M * 3) return;
    const long long row = e / 3;
    const int c = (int)(e - row * 3);
    const float x = pos[e];
    const float *g = gout + row * NE_W;
    float acc = g[6 * NE_FREQ + c];
#pragma unroll
    for (int i = 0; i < NE_FREQ; ++i) {
        const float s = nerf_scale(i);
        const float arg = __fmul_rn(s, x);
        acc += __fmul_rn(__fmul_rn(g[6 * i + c], cosf(arg)), s);
        acc += __fmul_rn(__fmul_rn(g[6 * i + 3 + c], -sinf(arg)), s);
    }
    gpos[e] = acc;
}

}  // namespace ptk

using namespace ptk;

extern "C" int ptk_nerf_embed_fwd(const float *positions, int64_t M, float *out, ptk_stream_t stream) {
    PTK_REQUIRE(M >= 0 && M < (1LL << 31), PTK_ERR_SHAPE, "nerf_embed_fwd: bad M = %lld", (long long)M);
    if (M == 0) return PTK_OK;
    PTK_REQUIRE(positions && out, PTK_ERR_SHAPE, "nerf_embed_fwd: null pointer");
    nerf_embed_fwd_kernel<<<(unsigned)ceil_div(M * NE_W, 256), 256, 0, as_stream(stream)>>>(positions, (long long)M, out);
    PTK_CHECK_LAUNCH();
    return PTK_OK;
}

extern "C" int ptk_nerf_embed_bwd(const float *positions, const float *grad_out, int64_t M, float *grad_positions,
                                  ptk_stream_t stream) {
    PTK_REQUIRE(M >= 0 && M < (1LL << 31), PTK_ERR_SHAPE, "nerf_embed_bwd: bad M = %lld", (long long)M);
    if (M == 0) return PTK_OK;
    PTK_REQUIRE(positions && grad_out && grad_positions, PTK_ERR_SHAPE, "nerf_embed_bwd: null pointer");
    nerf_embed_bwd_kernel<<<(unsigned)ceil_div(M * 3, 256), 256, 0, as_stream(stream)>>>(positions, grad_out, (long long)M,
                                                                                        grad_positions);
    PTK_CHECK_LAUNCH();
    return PTK_OK;
}
