"""Torch / numpy restatement (oracle tier O2) of the reference hot path -- TEST INFRASTRUCTURE.

An independent semantic cross-check of oracle/ptk_oracle.c: the same functions written with the
reference's own stack (eager torch ops), following the reference line by line.  Only tests/,
smoke() and bench.py's CPU-baseline legs may import it.

References (relative to /root/reference):
  pterotactyl/utility/utils.py:47-52    normalize_adj
  pterotactyl/utility/utils.py:134-148  calc_adj
  pterotactyl/utility/utils.py:75-130   adj_fuse_touch
  pterotactyl/utility/utils.py:152-187  batch_sample
  pterotactyl/utility/utils.py:204-217  chamfer_distance
  pterotactyl/reconstruction/vision/model.py:351-363  GCN_layer.forward
  pterotactyl/reconstruction/vision/model.py:381-397  Positional_Encoder.nerf_embedding (+ cat)
  PyTorch3D v0.5.0 (not vendored): loss/chamfer.py, ops/knn.py, ops/mesh_face_areas_normals.py,
  ops/sample_points_from_meshes.py::_rand_barycentric_coords.
"""
import numpy as np
import torch


def ieee_sqrt(t):
    """Correctly rounded fp32 sqrt.  torch's vectorised CPU sqrt is NOT always correctly rounded
    (measured here: 10 of 1088 face areas of test_objects/0.obj differ in the last bit from
    numpy / C sqrtf), whereas CUDA sqrtf -- what the reference executes -- is IEEE sqrt.rn.f32."""
    if t.is_cuda:
        return torch.sqrt(t)
    return torch.from_numpy(np.sqrt(t.detach().numpy()))


# ---------------------------------------------------------------- PyTorch3D pieces restated
def knn1(p1, p2, chunk=512):
    """knn_points(p1, p2, K=1): squared L2, lowest index on ties.  (B,P1,3),(B,P2,3) ->
    dists (B,P1) f32, idx (B,P1) i64.  Non-FMA arithmetic ((dx^2 + dy^2) + dz^2)."""
    B, P1, _ = p1.shape
    dists = torch.empty(B, P1, dtype=p1.dtype, device=p1.device)
    idx = torch.empty(B, P1, dtype=torch.int64, device=p1.device)
    for b in range(B):
        for s in range(0, P1, chunk):
            a = p1[b, s:s + chunk]
            diff = a[:, None, :] - p2[b][None, :, :]
            sq = diff * diff
            d = (sq[..., 0] + sq[..., 1]) + sq[..., 2]
            m, i = d.min(dim=1)  # torch.min returns the first minimal index on CPU
            if d.is_cuda:  # ... but any of several equal minima on CUDA: take the first one explicitly
                i = (d == m[:, None]).to(torch.uint8).argmax(dim=1)
            dists[b, s:s + chunk] = m
            idx[b, s:s + chunk] = i
    return dists, idx


def chamfer_distance(x, y):
    """pytorch3d.loss.chamfer_distance(x, y, batch_reduction=None) -> (cham (B,), None)."""
    dx, ix = knn1(x, y)
    dy, iy = knn1(y, x)
    cham = dx.sum(1) / x.shape[1] + dy.sum(1) / y.shape[1]
    return cham, ix, iy


def chamfer_autograd(x, y):
    """Differentiable Chamfer through gathers (the gradient knn_points_backward defines)."""
    with torch.no_grad():
        _, ix = knn1(x, y)
        _, iy = knn1(y, x)
    B = x.shape[0]
    ar = torch.arange(B, device=x.device)[:, None]
    dx = ((x - y[ar, ix]) ** 2).sum(-1)
    dy = ((y - x[ar, iy]) ** 2).sum(-1)
    return dx.sum(1) / x.shape[1] + dy.sum(1) / y.shape[1]


def mesh_face_areas(verts, faces):
    """mesh_face_areas_normals(V, F)[0] for a batch sharing faces: (B,V,3),(F,3) -> (B,F)."""
    v0 = verts[:, faces[:, 0]]
    v1 = verts[:, faces[:, 1]]
    v2 = verts[:, faces[:, 2]]
    a = v1 - v0
    b = v2 - v0
    cx = a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1]
    cy = a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2]
    cz = a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]
    n2 = (cx * cx + cy * cy) + cz * cz
    return ieee_sqrt(n2) * 0.5


def rand_barycentric_from(uv):
    """_rand_barycentric_coords with the uniforms passed in: uv (2,B,S)."""
    u, v = uv[0], uv[1]
    r = ieee_sqrt(u)
    return 1.0 - r, r * (1.0 - v), r * v


def face_cumweights(areas):
    """Integer cumulative weights (see ptk_oracle.c::orc_face_cumweights), vectorised numpy."""
    a = np.asarray(areas, np.float32).copy()
    a[np.isnan(a)] = 0.0
    B, F = a.shape
    cum = np.empty((B, F), np.uint64)
    for b in range(B):
        amax = a[b].max() if F else 0.0
        if amax == 0.0:
            q = np.ones(F, np.uint64)
        elif np.isinf(amax):
            q = np.isinf(a[b]).astype(np.uint64)
        else:
            q = ((a[b].astype(np.float64) / np.float64(amax)) * 4294967296.0).astype(np.uint64)
        cum[b] = np.cumsum(q, dtype=np.uint64)
    return cum


def pick_faces(cum, u_face):
    """(B,F) uint64, (B,S) f32 -> (B,S) int64 face index: first f with cum[f] > (t*total)>>24."""
    B, S = u_face.shape
    out = np.empty((B, S), np.int64)
    for b in range(B):
        total = int(cum[b, -1])
        t = np.clip((np.asarray(u_face[b], np.float32) * np.float32(16777216.0)).astype(np.int64), 0,
                    16777215)
        r = np.array([(int(tt) * total) >> 24 for tt in t], dtype=np.uint64)
        out[b] = np.searchsorted(cum[b], r, side="right")
    return out


def batch_sample(verts, faces, u_face, uv):
    """utils.batch_sample (utils.py:152-187) with explicit uniforms; differentiable w.r.t. verts."""
    with torch.no_grad():
        areas = mesh_face_areas(verts, faces)
        cum = face_cumweights(areas.cpu().numpy())
        fidx = torch.from_numpy(pick_faces(cum, u_face.cpu().numpy())).to(verts.device)
    B = verts.shape[0]
    ar = torch.arange(B, device=verts.device)[:, None]
    tri = faces[fidx]  # (B,S,3)
    A = verts[ar, tri[..., 0]]
    Bv = verts[ar, tri[..., 1]]
    Cv = verts[ar, tri[..., 2]]
    w0, w1, w2 = rand_barycentric_from(uv)
    pts = w0[:, :, None] * A + w1[:, :, None] * Bv + w2[:, :, None] * Cv
    return pts, fidx


# ---------------------------------------------------------------- adjacency (dense, as the reference)
def calc_adj(faces):
    n = int(faces.max()) + 1
    adj = torch.eye(n)
    v1, v2, v3 = faces[:, 0], faces[:, 1], faces[:, 2]
    for a, b in ((v1, v2), (v1, v3), (v2, v1), (v2, v3), (v3, v1), (v3, v2)):
        adj[(a, b)] = 1
    return adj


def normalize_adj(mx):
    rowsum = mx.sum(1)
    r_inv = (1.0 / rowsum).view(-1)
    r_inv[r_inv != r_inv] = 0.0
    return torch.mm(torch.eye(r_inv.shape[0]) * r_inv, mx)


def adj_fuse_touch(verts, faces, adj, sheet_verts, sheet_faces, num_grasps, finger, use_touch=True):
    vnp = verts.numpy()
    groups = {}
    for e, v in enumerate(vnp):
        groups.setdefault(v.tobytes(), []).append(e)
    central = []
    if use_touch:
        sheet_adj = calc_adj(sheet_faces)
        ns = sheet_adj.shape[0]
        n0 = adj.shape[0]
        k = (1 if finger else 4) * num_grasps
        central = [4 + i * ns + n0 for i in range(k)]
        new_adj = torch.zeros(n0 + k * ns, n0 + k * ns)
        new_adj[:n0, :n0] = adj
        for i in range(k):
            s = n0 + ns * i
            new_adj[s:s + ns, s:s + ns] = sheet_adj
        adj = new_adj
        all_faces = [faces] + [sheet_faces + verts.shape[0] + i * sheet_verts.shape[0] for i in range(k)]
        faces = torch.cat(all_faces)
    for cur in groups.values():
        if len(cur) > 1:
            for a in cur:
                for b in cur:
                    adj[a, b] = 1
                for c in central:
                    adj[a, c] = 1
                    adj[c, a] = 1
    return adj, faces


def dense_to_csr(adj):
    """Row-normalised dense adjacency -> (rowptr, col) int32; asserts every value is 1/deg."""
    a = adj.numpy()
    n = a.shape[0]
    rowptr = np.zeros(n + 1, np.int32)
    cols = []
    for i in range(n):
        nz = np.nonzero(a[i])[0]
        if len(nz):
            w = np.float32(1.0) / np.float32(len(nz))
            assert np.all(a[i, nz] == w), "adjacency row is not uniform 1/deg"
        cols.append(nz.astype(np.int32))
        rowptr[i + 1] = rowptr[i] + len(nz)
    return rowptr, np.concatenate(cols) if cols else np.zeros(0, np.int32)


# ---------------------------------------------------------------- GCN layer, dense (as the reference)
def gcn_layer_dense(features, weight, bias, adj, cut, do_cut, relu):
    h = torch.matmul(features, weight)
    if do_cut:
        length = round(h.shape[-1] * cut)
        out = torch.matmul(adj, h[:, :, :length])
        out = torch.cat((out, h[:, :, length:]), dim=-1)
        out[:, :, :length] += bias[:length]
    else:
        out = torch.matmul(adj, h) + bias
    return torch.relu(out) if relu else out


def gcn_dense(features, weights, biases, adj, cut):
    n = len(weights)
    for i in range(n):
        features = gcn_layer_dense(features, weights[i], biases[i], adj, cut, i < n - 1, i < n - 1)
    return features


# ---------------------------------------------------------------- positional embedding (as the reference)
def nerf_embedding(points):
    """Positional_Encoder.nerf_embedding followed by the cat with the positions
    (pterotactyl/reconstruction/vision/model.py:381-391, 396-397): (M,3) -> (M,63)."""
    parts = []
    for i in range(10):
        s = np.pi if i == 0 else np.pi * 2 * i
        parts += [torch.sin(s * points), torch.cos(s * points)]
    return torch.cat((torch.cat(parts, dim=-1), points), dim=-1)
