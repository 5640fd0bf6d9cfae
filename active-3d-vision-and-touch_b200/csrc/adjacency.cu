// GCN adjacency built on the device, emitted directly as CSR.
//
// Replaces the dense (Nv,Nv) fp32 construction of pterotactyl/utility/utils.py:
//   calc_adj        :134-148  identity + the 6 directed edges of every face
//   adj_fuse_touch  :75-130   block-diagonal vision + touch charts; vertices with the same 3-D position (byte
//                             equality, :81-86) are linked to each other and, both ways, to every touch-chart
//                             centre (:119-128 -- a Python double loop over the groups)
//   normalize_adj   :47-52    every entry of row i becomes fl32(1 / deg_i)
//
// Form: one bit per (row, col) in a workspace bitmap (Nv x ceil(Nv/32) words: 675 KB for the 2324-vertex graph,
// 32x smaller than the reference's dense fp32 matrix and never read by the host).  Marking is idempotent
// (atomicOr), so duplicate edges need no sort/unique; rows come out with ascending column ids because the
// bitmap is walked in order.  Two calls because the caller has to size `col` from rowptr[Nv]:
//   ptk_adj_count  memset -> mark faces + identity -> mark equal-position groups + centres -> degrees -> scan
//   ptk_adj_emit   bitmap rows -> col / val (= 1/deg_row) / val_t (= 1/deg_col: values of the transposed CSR;
//                  the pattern is symmetric by construction, so rowptr and col serve both directions)
// One-off set-up work (microseconds); the kernels are sized for clarity, not for a roofline.
#include "ptk_common.cuh"

namespace ptk {

constexpr int64_t ADJ_MAX_N = 65536;  // bitmap = 512 MB at the limit

struct AdjWs {
    uint32_t *bits;  // (n, wpr)
    int32_t *deg;    // (n)
    int wpr;
};

__host__ inline size_t adj_ws_bytes(int64_t n) {
    const int64_t wpr = (n + 31) / 32;
    return (size_t)(n * wpr * 4 + ((n * 4 + 255) / 256) * 256);
}

__host__ inline AdjWs adj_carve(void *ws, int64_t n) {
    AdjWs w;
    w.wpr = (int)((n + 31) / 32);
    w.deg = reinterpret_cast<int32_t *>(ws);
    w.bits = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(ws) + ((n * 4 + 255) / 256) * 256);
    return w;
}

__device__ __forceinline__ void set_bit(uint32_t *bits, int wpr, int r, int c) {
    atomicOr(bits + (size_t)r * wpr + (c >> 5), 1u << (c & 31));
}

// threads [0, F): the 6 directed edges of a face (utils.py:141-146); threads [F, F+n): the identity (:139)
__global__ void adj_mark_faces_kernel(const int32_t *__restrict__ faces, int F, int n, uint32_t *bits, int wpr) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < F) {
        const int a = faces[3 * t], b = faces[3 * t + 1], c = faces[3 * t + 2];
        if ((unsigned)a >= (unsigned)n || (unsigned)b >= (unsigned)n || (unsigned)c >= (unsigned)n) return;  // host checks
        set_bit(bits, wpr, a, b);
        set_bit(bits, wpr, a, c);
        set_bit(bits, wpr, b, a);
        set_bit(bits, wpr, b, c);
        set_bit(bits, wpr, c, a);
        set_bit(bits, wpr, c, b);
    } else if (t < F + n) {
        set_bit(bits, wpr, t - F, t - F);
    }
}

// Thread i scans every position j (shared-memory tiles) for a byte-identical (x,y,z): such pairs are linked, and a
// vertex that has at least one twin is linked both ways to every centre (utils.py:119-128).
constexpr int DUP_T = 256;
__global__ void __launch_bounds__(DUP_T)
adj_mark_twins_kernel(const uint32_t *__restrict__ pos, int n_pos, const int32_t *__restrict__ centres, int n_centres,
                      int n, uint32_t *bits, int wpr) {
    __shared__ uint32_t sx[DUP_T], sy[DUP_T], sz[DUP_T];
    const int i = blockIdx.x * DUP_T + threadIdx.x;
    const bool live = i < n_pos;
    uint32_t x = 0, y = 0, z = 0;
    if (live) {
        x = pos[3 * i];
        y = pos[3 * i + 1];
        z = pos[3 * i + 2];
    }
    bool twin = false;
    for (int j0 = 0; j0 < n_pos; j0 += DUP_T) {
        const int j = j0 + threadIdx.x;
        if (j < n_pos) {
            sx[threadIdx.x] = pos[3 * j];
            sy[threadIdx.x] = pos[3 * j + 1];
            sz[threadIdx.x] = pos[3 * j + 2];
        }
        __syncthreads();
        const int lim = min(DUP_T, n_pos - j0);
        if (live) {
            for (int k = 0; k < lim; ++k) {
                if (sx[k] == x && sy[k] == y && sz[k] == z && j0 + k != i) {
                    twin = true;
                    set_bit(bits, wpr, i, j0 + k);
                }
            }
        }
        __syncthreads();
    }
    if (live && twin) {
        for (int k = 0; k < n_centres; ++k) {
            const int c = centres[k];
            if ((unsigned)c >= (unsigned)n) continue;
            set_bit(bits, wpr, i, c);
            set_bit(bits, wpr, c, i);
        }
    }
}

// warp per row: degree = number of set bits
__global__ void adj_degree_kernel(const uint32_t *__restrict__ bits, int wpr, int n, int32_t *__restrict__ deg) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= n) return;
    int s = 0;
    for (int w = lane; w < wpr; w += 32) s += __popc(bits[(size_t)row * wpr + w]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) deg[row] = s;
}

// one CTA: rowptr = exclusive scan of deg (each thread owns a contiguous chunk)
constexpr int SCAN_T = 1024;
__global__ void __launch_bounds__(SCAN_T) adj_scan_kernel(const int32_t *__restrict__ deg, int n, int32_t *__restrict__ rowptr) {
    __shared__ int part[SCAN_T];
    const int per = (n + SCAN_T - 1) / SCAN_T;
    const int lo = min(n, (int)threadIdx.x * per), hi = min(n, lo + per);
    int s = 0;
    for (int i = lo; i < hi; ++i) s += deg[i];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int o = 1; o < SCAN_T; o <<= 1) {  // Hillis-Steele inclusive scan of the chunk sums
        const int v = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    int run = part[threadIdx.x] - s;
    for (int i = lo; i < hi; ++i) {
        rowptr[i] = run;
        run += deg[i];
    }
    if (threadIdx.x == SCAN_T - 1) rowptr[n] = part[SCAN_T - 1];
}

// warp per row: walk the row's words 32 at a time; a lane writes the set bits of its word behind the bits of the
// lower lanes => ascending column ids
__global__ void adj_emit_kernel(const uint32_t *__restrict__ bits, int wpr, int n, const int32_t *__restrict__ deg,
                                const int32_t *__restrict__ rowptr, int32_t *__restrict__ col, float *__restrict__ val,
                                float *__restrict__ val_t) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= n) return;
    const float w_row = __fdiv_rn(1.0f, (float)deg[row]);  // fl32(1 / rowsum), utils.py:49
    int base = rowptr[row];
    for (int w0 = 0; w0 < wpr; w0 += 32) {
        const int w = w0 + lane;
        uint32_t word = w < wpr ? bits[(size_t)row * wpr + w] : 0u;
        const int cnt = __popc(word);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        int at = base + incl - cnt;
        while (word) {
            const int b = __ffs(word) - 1;
            word &= word - 1;
            const int c = w * 32 + b;
            col[at] = c;
            if (val) val[at] = w_row;
            if (val_t) val_t[at] = __fdiv_rn(1.0f, (float)deg[c]);
            ++at;
        }
        base += __shfl_sync(0xffffffffu, incl, 31);
    }
}

}  // namespace ptk

using namespace ptk;

extern "C" size_t ptk_adj_workspace_bytes(int64_t n) {
    if (n <= 0 || n > ADJ_MAX_N) return 0;
    return adj_ws_bytes(n);
}

extern "C" int ptk_adj_count(const int32_t *faces, int64_t F, int64_t n, const float *positions, int64_t n_pos,
                             const int32_t *centres, int64_t n_centres, int32_t *rowptr, void *workspace,
                             size_t workspace_bytes, ptk_stream_t stream) {
    PTK_REQUIRE(n > 0 && n <= ADJ_MAX_N, PTK_ERR_SHAPE, "adj_count: n = %lld outside [1, %lld]", (long long)n,
                (long long)ADJ_MAX_N);
    PTK_REQUIRE(F >= 0 && F < (1LL << 30) && (F == 0 || faces), PTK_ERR_SHAPE, "adj_count: bad faces (F=%lld)", (long long)F);
    PTK_REQUIRE(n_pos >= 0 && n_pos <= n && (n_pos == 0 || positions), PTK_ERR_SHAPE,
                "adj_count: n_pos = %lld must lie in [0, n = %lld]", (long long)n_pos, (long long)n);
    PTK_REQUIRE(n_centres >= 0 && (n_centres == 0 || centres), PTK_ERR_SHAPE, "adj_count: bad centres");
    PTK_REQUIRE(rowptr && workspace, PTK_ERR_SHAPE, "adj_count: null pointer");
    PTK_REQUIRE(workspace_bytes >= adj_ws_bytes(n), PTK_ERR_WORKSPACE, "adj_count: workspace too small (%zu < %zu)",
                workspace_bytes, adj_ws_bytes(n));
    cudaStream_t st = as_stream(stream);
    const AdjWs w = adj_carve(workspace, n);
    PTK_CHECK_CUDA(cudaMemsetAsync(w.bits, 0, (size_t)n * w.wpr * 4, st));
    adj_mark_faces_kernel<<<(unsigned)ceil_div(F + n, 256), 256, 0, st>>>(faces, (int)F, (int)n, w.bits, w.wpr);
    PTK_CHECK_LAUNCH();
    if (n_pos > 1) {
        adj_mark_twins_kernel<<<(unsigned)ceil_div(n_pos, DUP_T), DUP_T, 0, st>>>(
            reinterpret_cast<const uint32_t *>(positions), (int)n_pos, centres, (int)n_centres, (int)n, w.bits, w.wpr);
        PTK_CHECK_LAUNCH();
    }
    adj_degree_kernel<<<(unsigned)ceil_div(n * 32, 256), 256, 0, st>>>(w.bits, w.wpr, (int)n, w.deg);
    PTK_CHECK_LAUNCH();
    adj_scan_kernel<<<1, SCAN_T, 0, st>>>(w.deg, (int)n, rowptr);
    PTK_CHECK_LAUNCH();
    return PTK_OK;
}

extern "C" int ptk_adj_emit(int64_t n, const int32_t *rowptr, int32_t *col, float *val, float *val_t,
                            const void *workspace, size_t workspace_bytes, ptk_stream_t stream) {
    PTK_REQUIRE(n > 0 && n <= ADJ_MAX_N, PTK_ERR_SHAPE, "adj_emit: n = %lld outside [1, %lld]", (long long)n,
                (long long)ADJ_MAX_N);
    PTK_REQUIRE(rowptr && col && workspace, PTK_ERR_SHAPE, "adj_emit: null pointer");
    PTK_REQUIRE(workspace_bytes >= adj_ws_bytes(n), PTK_ERR_WORKSPACE, "adj_emit: workspace too small");
    const AdjWs w = adj_carve(const_cast<void *>(workspace), n);
    adj_emit_kernel<<<(unsigned)ceil_div(n * 32, 256), 256, 0, as_stream(stream)>>>(w.bits, w.wpr, (int)n, w.deg, rowptr,
                                                                                   col, val, val_t);
    PTK_CHECK_LAUNCH();
    return PTK_OK;
}
