"""Minimal Wavefront OBJ reader / writer with the call shapes of pytorch3d.io.obj_io that the
reference uses (pterotactyl/utility/utils.py:23,195,362,375,600):

    verts, faces, aux = load_obj(path);  faces.verts_idx  -> (F,3) int64, zero based
    save_obj(path, verts, faces, decimal_places=None)

Polygons are fan-triangulated; v/vt/vn triples keep the first (vertex) index; negative indices
are relative to the vertices read so far (OBJ convention).
"""
from collections import namedtuple

import numpy as np
import torch

Faces = namedtuple("Faces", "verts_idx normals_idx textures_idx materials_idx")
Properties = namedtuple("Properties", "normals verts_uvs material_colors texture_images texture_atlas")


def load_obj(f, load_textures=False, **_unused):
    verts, tris = [], []
    opened = isinstance(f, (str, bytes)) or hasattr(f, "__fspath__")
    fh = open(f, "r") if opened else f
    try:
        for line in fh:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "v" and len(tok) >= 4:
                verts.append((float(tok[1]), float(tok[2]), float(tok[3])))
            elif tok[0] == "f":
                idx = []
                for t in tok[1:]:
                    i = int(t.split("/")[0])
                    idx.append(i - 1 if i > 0 else len(verts) + i)
                for k in range(1, len(idx) - 1):
                    tris.append((idx[0], idx[k], idx[k + 1]))
    finally:
        if opened:
            fh.close()
    v = torch.from_numpy(np.asarray(verts, np.float32).reshape(-1, 3))
    t = torch.from_numpy(np.asarray(tris, np.int64).reshape(-1, 3))
    if len(tris) and (int(t.max()) >= len(verts) or int(t.min()) < 0):
        raise ValueError("Faces have invalid indices")
    empty = torch.zeros((0, 3), dtype=torch.int64)
    faces = Faces(verts_idx=t, normals_idx=empty, textures_idx=empty, materials_idx=torch.zeros((0,), dtype=torch.int64))
    return v, faces, Properties(None, None, None, None, None)


def save_obj(f, verts, faces=None, decimal_places=None):
    if verts.dim() != 2 or verts.shape[1] != 3:
        raise ValueError("Argument 'verts' should either be empty or of shape (num_verts, 3).")
    if faces is not None and faces.numel() and (faces.dim() != 2 or faces.shape[1] != 3):
        raise ValueError("Argument 'faces' should either be empty or of shape (num_faces, 3).")
    fmt = "%f" if decimal_places is None else "%." + str(int(decimal_places)) + "f"
    v = verts.detach().cpu().numpy()
    lines = ["v " + " ".join(fmt % c for c in row) for row in v]
    if faces is not None and faces.numel():
        fa = faces.detach().cpu().numpy() + 1
        lines += ["f %d %d %d" % tuple(row) for row in fa]
    text = "\n".join(lines) + ("\n" if lines else "")
    if isinstance(f, (str, bytes)) or hasattr(f, "__fspath__"):
        with open(f, "w") as fh:
            fh.write(text)
    else:
        f.write(text)
