// Host-buffer entry points of libptk_b200 (include/ptk.h: ptk_host_*).
//
// The end-to-end path a non-torch caller binds: HOST pointers in, HOST pointers out; the call does
// the H2D copies, launches the same kernels as the device-pointer API on its own stream, copies the
// results back and synchronises.  Device buffers live in the context and only ever grow.
#include <stdlib.h>

#include "ptk_common.cuh"

struct ptk_host_ctx {
    int device;
    cudaStream_t stream;       // kernels + result copies
    cudaStream_t copy_stream;  // input copies, so that chunk i+1 uploads while chunk i computes
    cudaEvent_t ev_in[9];
    void *buf[16];
    size_t cap[16];
};

namespace ptk {

enum { HB_X = 0, HB_Y, HB_WS, HB_IDXX, HB_IDXY, HB_CHAM, HB_GCHAM, HB_GX, HB_GY, HB_VERTS, HB_FACES,
       HB_UF, HB_UV, HB_FIDX, HB_SWS, HB_ACC };

static int ensure(ptk_host_ctx *c, int slot, size_t bytes) {
    if (bytes <= c->cap[slot]) return PTK_OK;
    if (c->buf[slot]) PTK_CHECK_CUDA(cudaFree(c->buf[slot]));
    c->buf[slot] = nullptr;
    c->cap[slot] = 0;
    PTK_CHECK_CUDA(cudaMalloc(&c->buf[slot], bytes));
    c->cap[slot] = bytes;
    return PTK_OK;
}

// acc[i] = (first ? 0 : acc[i]) + s * v[i]
__global__ void axpy_kernel(float *__restrict__ acc, const float *__restrict__ v, float s, long long n,
                            int first) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) acc[i] = (first ? 0.f : acc[i]) + s * v[i];
}
__global__ void scale_kernel(float *__restrict__ out, const float *__restrict__ v, float s, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = s * v[i];
}

}  // namespace ptk

using namespace ptk;

#define PTK_TRY(expr)          \
    do {                       \
        int _rc = (expr);      \
        if (_rc) return _rc;   \
    } while (0)

extern "C" ptk_host_ctx *ptk_host_ctx_create(int device) {
    ptk_host_ctx *c = (ptk_host_ctx *)calloc(1, sizeof(ptk_host_ctx));
    if (!c) return nullptr;
    c->device = device;
    bool ok = cudaSetDevice(device) == cudaSuccess &&
              cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; ok && i < 9; ++i) ok = cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
        set_error("host_ctx_create: cannot open device %d: %s", device,
                  cudaGetErrorString(cudaGetLastError()));
        free(c);
        return nullptr;
    }
    return c;
}

extern "C" void ptk_host_ctx_destroy(ptk_host_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (int i = 0; i < 16; ++i)
        if (c->buf[i]) cudaFree(c->buf[i]);
    for (int i = 0; i < 9; ++i)
        if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]);
    cudaStreamDestroy(c->copy_stream);
    cudaStreamDestroy(c->stream);
    free(c);
}

namespace ptk {
int chamfer_resolve_algo(int64_t P1, int64_t P2);  // chamfer.cu: the scan ptk_chamfer_fwd will run for these sizes
}

extern "C" int ptk_host_chamfer(ptk_host_ctx *c, const float *x, const float *y, int64_t B, int64_t P1,
                                int64_t P2, float *cham, int32_t *idx_x, int32_t *idx_y,
                                const float *grad_cham, float *grad_x, float *grad_y) {
    PTK_NVTX("ptk_host_chamfer");
    PTK_REQUIRE(c && x && y && cham, PTK_ERR_SHAPE, "host_chamfer: null pointer");
    PTK_REQUIRE(B > 0 && P1 > 0 && P2 > 0, PTK_ERR_SHAPE, "host_chamfer: empty input");
    PTK_CHECK_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const size_t nx = (size_t)B * P1, ny = (size_t)B * P2;
    PTK_TRY(ensure(c, HB_X, nx * 12));
    PTK_TRY(ensure(c, HB_Y, ny * 12));
    PTK_TRY(ensure(c, HB_WS, ptk_chamfer_workspace_bytes(B, P1, P2)));
    PTK_TRY(ensure(c, HB_IDXX, nx * 4));
    PTK_TRY(ensure(c, HB_IDXY, ny * 4));
    PTK_TRY(ensure(c, HB_CHAM, (size_t)B * 4));
    float *dx = (float *)c->buf[HB_X], *dy = (float *)c->buf[HB_Y];
    int32_t *dix = (int32_t *)c->buf[HB_IDXX], *diy = (int32_t *)c->buf[HB_IDXY];
    float *dcham = (float *)c->buf[HB_CHAM];
    float *dgc = nullptr, *dgx = nullptr, *dgy = nullptr;
    if (grad_cham) {
        // backward always runs on the device when grad_cham is given; the gradients are copied back
        // only where a host pointer is supplied (a trainer keeps them on the device)
        PTK_TRY(ensure(c, HB_GCHAM, (size_t)B * 4));
        PTK_TRY(ensure(c, HB_GX, nx * 12));
        PTK_TRY(ensure(c, HB_GY, ny * 12));
        dgc = (float *)c->buf[HB_GCHAM];
        dgx = (float *)c->buf[HB_GX];
        dgy = (float *)c->buf[HB_GY];
        PTK_CHECK_CUDA(cudaMemcpyAsync(dgc, grad_cham, (size_t)B * 4, cudaMemcpyHostToDevice, st));
    }
    // Chunked pipeline over whole cloud pairs: the upload of chunk i+1 (copy stream) overlaps the kernels of
    // chunk i.  The first chunk is small (B/8) so that little of its upload is exposed; the others keep enough
    // CTAs for >= 4 waves (see plan_nn in chamfer.cu) so the scan itself does not change; at most 8 chunks.
    const int64_t Pm = P1 > P2 ? P1 : P2;
    const int64_t ctas_per_pair = 2 * ceil_div(Pm, 1024);
    int64_t bc_min = ceil_div(4LL * sm_count() * 4, ctas_per_pair);
    if (bc_min < 1) bc_min = 1;
    int64_t bounds[10];
    int nchunks = 0;
    bounds[0] = 0;
    if (chamfer_resolve_algo(P1, P2) == PTK_CHAMFER_PRUNED && B >= 8) {
        // the pruned scan is faster than the upload of its own input: equal chunks, the copy engine never waits and
        // only the last chunk's kernels are exposed
        for (int ci = 1; ci <= 8; ++ci) bounds[++nchunks] = B * ci / 8;
    } else if (B >= 2 * bc_min) {
        const int64_t first = B / 8 > 0 ? B / 8 : 1;
        int64_t rest_chunks = (B - first) / bc_min;
        if (rest_chunks < 1) rest_chunks = 1;
        if (rest_chunks > 7) rest_chunks = 7;
        const int64_t bc = ceil_div(B - first, rest_chunks);
        bounds[++nchunks] = first;
        while (bounds[nchunks] < B) {
            const int64_t nxt = bounds[nchunks] + bc < B ? bounds[nchunks] + bc : B;
            bounds[++nchunks] = nxt;
        }
    } else {
        bounds[++nchunks] = B;
    }
    for (int ci = 0; ci < nchunks; ++ci) {
        const int64_t b0 = bounds[ci], nb = bounds[ci + 1] - bounds[ci];
        const size_t ox = (size_t)b0 * P1, oy = (size_t)b0 * P2;
        PTK_CHECK_CUDA(cudaMemcpyAsync(dx + ox * 3, x + ox * 3, (size_t)nb * P1 * 12, cudaMemcpyHostToDevice, c->copy_stream));
        PTK_CHECK_CUDA(cudaMemcpyAsync(dy + oy * 3, y + oy * 3, (size_t)nb * P2 * 12, cudaMemcpyHostToDevice, c->copy_stream));
        PTK_CHECK_CUDA(cudaEventRecord(c->ev_in[ci], c->copy_stream));
        PTK_CHECK_CUDA(cudaStreamWaitEvent(st, c->ev_in[ci], 0));
        PTK_TRY(ptk_chamfer_fwd(dx + ox * 3, dy + oy * 3, nb, P1, P2, nullptr, dix + ox, nullptr, diy + oy, dcham + b0,
                                c->buf[HB_WS], c->cap[HB_WS], st));
        if (grad_cham)
            PTK_TRY(ptk_chamfer_bwd(dx + ox * 3, dy + oy * 3, dix + ox, diy + oy, dgc + b0, nb, P1, P2, dgx + ox * 3,
                                    dgy + oy * 3, st));
    }
    PTK_CHECK_CUDA(cudaMemcpyAsync(cham, dcham, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
    if (idx_x) PTK_CHECK_CUDA(cudaMemcpyAsync(idx_x, dix, nx * 4, cudaMemcpyDeviceToHost, st));
    if (idx_y) PTK_CHECK_CUDA(cudaMemcpyAsync(idx_y, diy, ny * 4, cudaMemcpyDeviceToHost, st));
    if (grad_cham && grad_x) PTK_CHECK_CUDA(cudaMemcpyAsync(grad_x, dgx, nx * 12, cudaMemcpyDeviceToHost, st));
    if (grad_cham && grad_y) PTK_CHECK_CUDA(cudaMemcpyAsync(grad_y, dgy, ny * 12, cudaMemcpyDeviceToHost, st));
    PTK_CHECK_CUDA(cudaStreamSynchronize(st));
    return PTK_OK;
}

extern "C" int ptk_host_mesh_chamfer(ptk_host_ctx *c, const float *verts, int64_t B, int64_t V,
                                     const int32_t *faces, int64_t F, const float *gt, int64_t P2,
                                     const float *u_face, const float *uv, int64_t S, int64_t repeat,
                                     float *cd, const float *grad_cd, float *grad_verts) {
    PTK_NVTX("ptk_host_mesh_chamfer");
    PTK_REQUIRE(c && verts && faces && gt && u_face && uv && cd, PTK_ERR_SHAPE, "host_mesh_chamfer: null pointer");
    PTK_REQUIRE(B > 0 && V > 0 && F > 0 && P2 > 0 && S > 0 && repeat > 0, PTK_ERR_SHAPE,
                "host_mesh_chamfer: empty input");
    PTK_CHECK_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const size_t ns = (size_t)B * S, ny = (size_t)B * P2, nv = (size_t)B * V * 3;
    const bool bwd = grad_cd && grad_verts;
    PTK_TRY(ensure(c, HB_VERTS, nv * 4));
    PTK_TRY(ensure(c, HB_FACES, (size_t)F * 12));
    PTK_TRY(ensure(c, HB_Y, ny * 12));
    PTK_TRY(ensure(c, HB_UF, ns * 4 * repeat));
    PTK_TRY(ensure(c, HB_UV, ns * 8 * repeat));
    PTK_TRY(ensure(c, HB_X, ns * 12));
    PTK_TRY(ensure(c, HB_FIDX, ns * 4));
    PTK_TRY(ensure(c, HB_SWS, ptk_sample_workspace_bytes(B, F)));
    PTK_TRY(ensure(c, HB_WS, ptk_chamfer_workspace_bytes(B, S, P2)));
    PTK_TRY(ensure(c, HB_IDXX, ns * 4));
    PTK_TRY(ensure(c, HB_IDXY, ny * 4));
    PTK_TRY(ensure(c, HB_CHAM, (size_t)B * 4));
    PTK_TRY(ensure(c, HB_ACC, (size_t)B * 4 + nv * 4));
    float *d_cd = (float *)c->buf[HB_ACC];
    float *d_gv_acc = d_cd + B;
    if (bwd) {
        PTK_TRY(ensure(c, HB_GCHAM, (size_t)B * 8));
        PTK_TRY(ensure(c, HB_GX, ns * 12));
        PTK_TRY(ensure(c, HB_GY, nv * 4));
    }
    PTK_CHECK_CUDA(cudaMemcpyAsync(c->buf[HB_VERTS], verts, nv * 4, cudaMemcpyHostToDevice, st));
    PTK_CHECK_CUDA(cudaMemcpyAsync(c->buf[HB_FACES], faces, (size_t)F * 12, cudaMemcpyHostToDevice, st));
    PTK_CHECK_CUDA(cudaMemcpyAsync(c->buf[HB_Y], gt, ny * 12, cudaMemcpyHostToDevice, st));
    PTK_CHECK_CUDA(cudaMemcpyAsync(c->buf[HB_UF], u_face, ns * 4 * repeat, cudaMemcpyHostToDevice, st));
    PTK_CHECK_CUDA(cudaMemcpyAsync(c->buf[HB_UV], uv, ns * 8 * repeat, cudaMemcpyHostToDevice, st));
    float *d_gc = nullptr;
    if (bwd) {
        d_gc = (float *)c->buf[HB_GCHAM];
        PTK_CHECK_CUDA(cudaMemcpyAsync(d_gc + B, grad_cd, (size_t)B * 4, cudaMemcpyHostToDevice, st));
        scale_kernel<<<(unsigned)ceil_div(B, 256), 256, 0, st>>>(d_gc, d_gc + B, 1.0f / (float)repeat, B);
        PTK_CHECK_LAUNCH();
    }
    for (int64_t r = 0; r < repeat; ++r) {
        const float *uf = (const float *)c->buf[HB_UF] + r * ns;
        const float *uvr = (const float *)c->buf[HB_UV] + r * 2 * ns;
        PTK_TRY(ptk_sample_fwd((float *)c->buf[HB_VERTS], B, V, (int32_t *)c->buf[HB_FACES], F, uf, uvr, S,
                               (float *)c->buf[HB_X], (int32_t *)c->buf[HB_FIDX], c->buf[HB_SWS],
                               c->cap[HB_SWS], st));
        PTK_TRY(ptk_chamfer_fwd((float *)c->buf[HB_X], (float *)c->buf[HB_Y], B, S, P2, nullptr,
                                (int32_t *)c->buf[HB_IDXX], nullptr, (int32_t *)c->buf[HB_IDXY],
                                (float *)c->buf[HB_CHAM], c->buf[HB_WS], c->cap[HB_WS], st));
        axpy_kernel<<<(unsigned)ceil_div(B, 256), 256, 0, st>>>(d_cd, (float *)c->buf[HB_CHAM],
                                                               1.0f / (float)repeat, B, r == 0);
        PTK_CHECK_LAUNCH();
        if (bwd) {
            PTK_TRY(ptk_chamfer_bwd((float *)c->buf[HB_X], (float *)c->buf[HB_Y], (int32_t *)c->buf[HB_IDXX],
                                    (int32_t *)c->buf[HB_IDXY], d_gc, B, S, P2, (float *)c->buf[HB_GX],
                                    nullptr, st));
            PTK_TRY(ptk_sample_bwd((float *)c->buf[HB_GX], (int32_t *)c->buf[HB_FIDX], uvr,
                                   (int32_t *)c->buf[HB_FACES], B, V, F, S, (float *)c->buf[HB_GY], st));
            axpy_kernel<<<(unsigned)ceil_div((int64_t)nv, 256), 256, 0, st>>>(
                d_gv_acc, (float *)c->buf[HB_GY], 1.0f, (long long)nv, r == 0);
            PTK_CHECK_LAUNCH();
        }
    }
    PTK_CHECK_CUDA(cudaMemcpyAsync(cd, d_cd, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
    if (bwd) PTK_CHECK_CUDA(cudaMemcpyAsync(grad_verts, d_gv_acc, nv * 4, cudaMemcpyDeviceToHost, st));
    PTK_CHECK_CUDA(cudaStreamSynchronize(st));
    return PTK_OK;
}
