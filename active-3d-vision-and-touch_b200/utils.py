"""Drop-in replacements for the hot-path functions of pterotactyl/utility/utils.py -- same names,
arguments, return values and error behaviour, backed by libptk_b200.so.

    chamfer_distance(verts, faces, gt_points, num=1000, repeat=3)   utils.py:204-217
    batch_sample(verts, faces, num=10000)                            utils.py:152-187
    calc_adj / normalize_adj / adj_fuse_touch / adj_init             utils.py:47-148
    load_mesh_touch / load_mesh_vision                               utils.py:30-36,194-200

`install()` (package __init__) patches these over an importable `pterotactyl.utility.utils` so the
reference's train / eval / policy scripts pick the new path up without edits.
"""
import os
import weakref

import numpy as np
import torch

from . import graph as _graph
from . import obj_io, ops

_faces_cache = {}  # id(faces) -> (weakref to that very tensor, its _version, int32 copy, largest vertex id)


def _faces_i32(faces, n_verts=None):
    """(F,3) int64 faces of the reference -> int32 copy on the same device, validated.

    The copy and the id range are cached PER LIVE TENSOR OBJECT: a hit requires that the weak reference still
    points at this very tensor and that its version counter is unchanged, so a new tensor that happens to reuse
    a freed tensor's address (the caching allocator does that all the time) can never be served a stale copy.
    Vertex ids are checked against [0, n_verts): torch's own indexing would raise where the kernels would read out
    of bounds (one host read per new faces tensor, none afterwards)."""
    key = id(faces)
    hit = _faces_cache.get(key)
    if hit is None or hit[0]() is not faces or hit[1] != faces._version:
        if faces.dim() != 2 or faces.shape[1] != 3:
            raise ValueError(f"faces must be (F,3), got {tuple(faces.shape)}")
        f32 = faces.contiguous() if faces.dtype == torch.int32 else faces.to(torch.int32).contiguous()
        lo, hi = (int(faces.min()), int(faces.max())) if faces.numel() else (0, -1)
        if lo < 0:
            raise IndexError(f"negative vertex id {lo} in faces")
        if len(_faces_cache) > 64:
            for k in [k for k, v in _faces_cache.items() if v[0]() is None]:
                del _faces_cache[k]
            if len(_faces_cache) > 64:
                _faces_cache.clear()
        try:
            ref = weakref.ref(faces)
        except TypeError:  # pragma: no cover
            return f32
        hit = _faces_cache[key] = (ref, faces._version, f32, hi)
    if n_verts is not None and hit[3] >= n_verts:
        raise IndexError(f"faces reference vertex {hit[3]} but the meshes have {n_verts} vertices")
    return hit[2]


def draw_uniforms(bs, num, device, generator=None):
    """The RNG stream of one batch_sample call in the explicit-uniform contract, in the reference's consumption
    order: first the face draw (in place of Tensor.multinomial, utils.py:170), then torch.rand(2, bs, num)
    (_rand_barycentric_coords, utils.py:179)."""
    u_face = torch.rand(bs, num, device=device, generator=generator)
    uv = torch.rand(2, bs, num, device=device, generator=generator)
    return u_face, uv


# How batch_sample draws the face of every sample when no explicit uniforms are passed:
#   "uniform"      one torch.rand(bs, num) + an order-independent integer prefix sum, fused in ptk_sample_fwd.
#                  Bit-exact for a given uniform tensor on any device / torch version; NOT the reference's stream.
#   "multinomial"  the reference's own statements (utils.py:165-170) on areas from ptk_mesh_face_areas: ATen's
#                  Tensor.multinomial draws the faces, so with the same seed the sampled points are bit-identical to
#                  the reference's batch_sample.  install() selects this mode: existing scripts keep their RNG stream.
face_draw = "uniform"


def batch_sample(verts, faces, num=10000, generator=None, uniforms=None, face_draw=None):
    """Sample `num` area-weighted surface points per mesh.  verts (B,V,3), faces (F,3) shared by the
    batch -> (B,num,3).  Gradients flow to verts only (utils.py:152-187)."""
    f32 = _faces_i32(faces, verts.shape[-2])
    mode = face_draw or globals()["face_draw"]
    if uniforms is None and mode == "multinomial":
        with torch.no_grad():
            Ar = ops.mesh_face_areas(verts, f32)                       # utils.py:163-164
            Ar[Ar != Ar] = 0                                           # utils.py:165
            Ar = torch.abs(Ar / Ar.sum(1).unsqueeze(1))                # utils.py:166
            Ar[Ar != Ar] = 1                                           # utils.py:167
            fidx = Ar.multinomial(num, replacement=True, generator=generator)   # utils.py:170 -- ATen draws
        uv = torch.rand(2, verts.shape[0], num, dtype=torch.float32, device=verts.device, generator=generator)  # :179
        pts, _ = ops.sample_points(verts, f32, None, uv, face_idx=fidx)
        return pts
    if mode not in ("uniform", "multinomial"):
        raise ValueError(f"face_draw must be 'uniform' or 'multinomial', got {mode!r}")
    u_face, uv = uniforms if uniforms is not None else draw_uniforms(verts.shape[0], num, verts.device, generator)
    pts, _ = ops.sample_points(verts, f32, u_face, uv)
    return pts


fused_mesh_chamfer = True  # chamfer_distance as one autograd node (ops.mesh_chamfer); tests flip this


def chamfer_distance(verts, faces, gt_points, num=1000, repeat=3, generator=None, uniforms=None, face_draw=None):
    """Chamfer distance between a predicted mesh and a ground-truth cloud: mean over `repeat`
    independent surface samplings of pytorch3d-style chamfer(pred_points, gt_points,
    batch_reduction=None).  Returns (B,) (utils.py:204-217).

    The random draws are made here, repeat by repeat, in the reference's consumption order (face draw, then
    torch.rand(2,bs,num): utils.py:170,179); sampling, Chamfer and the mean then run as ONE call / one autograd node
    (ptk_mesh_chamfer_fwd / bwd)."""
    R = max(int(repeat), 1)
    if not fused_mesh_chamfer:
        cds = []
        for r in range(R):
            uni = uniforms[r] if uniforms is not None else None
            pred_points = batch_sample(verts, faces, num=num, generator=generator, uniforms=uni, face_draw=face_draw)
            cd, _, _ = ops.chamfer(pred_points, gt_points)
            cds.append(cd)
        return cds[0] if len(cds) == 1 else torch.stack(cds).mean(dim=0)
    f32 = _faces_i32(faces, verts.shape[-2])
    mode = face_draw or globals()["face_draw"]
    if mode not in ("uniform", "multinomial"):
        raise ValueError(f"face_draw must be 'uniform' or 'multinomial', got {mode!r}")
    bs, dev = verts.shape[0], verts.device
    ufs, uvs, fis = [], [], []
    for r in range(R):
        if uniforms is not None:
            uf, uv = uniforms[r]
            ufs.append(uf)
            uvs.append(uv)
        elif mode == "multinomial":
            with torch.no_grad():
                Ar = ops.mesh_face_areas(verts, f32)                       # utils.py:163-164
                Ar[Ar != Ar] = 0                                           # utils.py:165
                Ar = torch.abs(Ar / Ar.sum(1).unsqueeze(1))                # utils.py:166
                Ar[Ar != Ar] = 1                                           # utils.py:167
                fis.append(Ar.multinomial(num, replacement=True, generator=generator))   # utils.py:170
            uvs.append(torch.rand(2, bs, num, dtype=torch.float32, device=dev, generator=generator))  # :179
        else:
            uf, uv = draw_uniforms(bs, num, dev, generator)
            ufs.append(uf)
            uvs.append(uv)
    uv = torch.stack(uvs)
    if fis:
        return ops.mesh_chamfer(verts, gt_points, f32, None, uv, face_idx=torch.stack(fis))
    return ops.mesh_chamfer(verts, gt_points, f32, torch.stack(ufs), uv)


# ------------------------------------------------------------------------------------- adjacency
def calc_adj(faces):
    """Dense binary adjacency (+ identity) from faces (utils.py:134-148)."""
    v1, v2, v3 = faces[:, 0], faces[:, 1], faces[:, 2]
    num_verts = int(faces.max())
    adj = torch.eye(num_verts + 1).to(faces.device)
    adj[(v1, v2)] = 1
    adj[(v1, v3)] = 1
    adj[(v2, v1)] = 1
    adj[(v2, v3)] = 1
    adj[(v3, v1)] = 1
    adj[(v3, v2)] = 1
    return adj


def normalize_adj(mx):
    """Row-normalise a binary adjacency (utils.py:47-52).  The result is registered with the CSR
    cache so GCN layers never re-scan the dense matrix."""
    rowsum = mx.sum(1)
    r_inv = (1.0 / rowsum).view(-1)
    r_inv[r_inv != r_inv] = 0.0
    out = mx * r_inv[:, None]  # == mm(eye * r_inv, mx): products with 0/1 are exact
    return out


def adj_fuse_touch(verts, faces, adj, args):
    """Fuse vision and touch charts into one graph (utils.py:75-130)."""
    vnp = verts.data.cpu().numpy()
    groups = {}
    for e, v in enumerate(vnp):
        groups.setdefault(v.tobytes(), []).append(e)
    central_points = []
    if args.use_touch:
        sheet_verts, sheet_faces = load_mesh_touch(_object_path("touch_chart.obj"), device=faces.device)
        sheet_adj = calc_adj(sheet_faces)
        ns, n0 = sheet_adj.shape[0], adj.shape[0]
        k = (1 if args.finger else 4) * args.num_grasps
        central_points = [4 + i * ns + n0 for i in range(k)]
        new_adj = torch.zeros((n0 + k * ns, n0 + k * ns), device=adj.device)
        new_adj[:n0, :n0] = adj
        for i in range(k):
            s = n0 + ns * i
            new_adj[s:s + ns, s:s + ns] = sheet_adj
        adj = new_adj
        all_faces = [faces] + [sheet_faces + verts.shape[0] + i * sheet_verts.shape[0] for i in range(k)]
        faces = torch.cat(all_faces)
    # vertices sharing a 3-D position talk to each other and to every touch-chart centre
    dup = [g for g in groups.values() if len(g) > 1]
    if dup:
        a = adj.cpu()
        for cur in dup:
            idx = torch.tensor(cur)
            a[idx[:, None], idx[None, :]] = 1
            if central_points:
                c = torch.tensor(central_points)
                a[idx[:, None], c[None, :]] = 1
                a[c[:, None], idx[None, :]] = 1
        adj = a.to(adj.device)
    return adj, faces


def _adj_init_device(verts, faces, args):
    """adj_init for CUDA inputs: both graphs are built on the device straight from the faces
    (Graph.from_faces -> ptk_adj_count / ptk_adj_emit); the dense tensors the reference's callers index
    adj_info with are scattered from the CSR and registered, so they are never scanned."""
    dev = faces.device
    n0 = int(faces.max()) + 1  # calc_adj: eye(faces.max() + 1)
    g0 = _graph.Graph.from_faces(faces, n0)
    adj_info = {}
    if getattr(args, "use_touch", False):
        sheet_verts, sheet_faces = load_mesh_touch(_object_path("touch_chart.obj"), device=dev)
        ns = int(sheet_faces.max()) + 1
        k = (1 if args.finger else 4) * args.num_grasps
        n = n0 + k * ns
        if verts.shape[0] > n:
            raise IndexError(f"{verts.shape[0]} vertex positions for a fused graph of {n} vertices")
        # adjacency blocks sit at n0 + i*ns (utils.py:100-106); the face list is offset by the vertex counts (:110-116)
        adj_faces = torch.cat([faces] + [sheet_faces + (n0 + i * ns) for i in range(k)])
        out_faces = torch.cat([faces] + [sheet_faces + (verts.shape[0] + i * sheet_verts.shape[0]) for i in range(k)])
        centres = [4 + i * ns + n0 for i in range(k)]
        g = _graph.Graph.from_faces(adj_faces, n, positions=verts, centres=centres)
    else:
        g, out_faces = g0, faces
    for key, gr in (("origional", g0), ("adj", g)):
        adj_info[key] = gr.dense()
        _graph.register(adj_info[key], gr)
    adj_info["faces"] = out_faces
    return adj_info


def adj_init(verts, faces, args):
    """{'origional', 'adj', 'faces'} (utils.py:56-71).  CUDA inputs take the device builder; CPU tensors follow
    the reference's dense construction literally (host-side tests and tools)."""
    if faces.is_cuda:
        return _adj_init_device(verts, faces, args)
    adj = calc_adj(faces)
    adj_info = {"origional": normalize_adj(adj.clone())}
    if args.use_touch:
        adj, faces = adj_fuse_touch(verts, faces, adj, args)
    adj_info["adj"] = normalize_adj(adj)
    adj_info["faces"] = faces
    if adj_info["adj"].is_cuda:
        for k in ("origional", "adj"):
            _graph.graph_of(adj_info[k])
    return adj_info


# ------------------------------------------------------------------------------------- mesh loading
_OBJECT_DIR = None


def set_object_dir(path):
    """Directory holding vision_charts.obj / touch_chart.obj (pterotactyl/objects)."""
    global _OBJECT_DIR
    _OBJECT_DIR = path


def _object_path(name):
    if _OBJECT_DIR is not None:
        return os.path.join(_OBJECT_DIR, name)
    try:
        import pterotactyl.objects as objects  # the reference's asset package (utils.py:25)
        return os.path.join(os.path.dirname(objects.__file__), name)
    except Exception as exc:  # pragma: no cover
        raise FileNotFoundError(
            f"cannot locate {name}: call set_object_dir(<pterotactyl/objects>) or make pterotactyl importable") from exc


def load_mesh_touch(obj, device="cuda"):
    """verts (V,3) f32, faces (F,3) i64 on the GPU (utils.py:194-200)."""
    verts, faces, _ = obj_io.load_obj(obj)
    return verts.float().to(device), faces.verts_idx.long().to(device)


def load_mesh_vision(args, obj, device="cuda"):
    """(adj_info, verts) for the vision charts (utils.py:30-36)."""
    verts, faces = load_mesh_touch(obj, device=device)
    return adj_init(verts, faces, args), verts
