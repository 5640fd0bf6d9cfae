"""pytorch3d.ops.mesh_face_areas_normals (utils.py:21,164)."""
import ptk_b200


def mesh_face_areas_normals(verts, faces):
    if verts.dim() != 2 or verts.shape[1] != 3:
        raise ValueError("verts need to be of shape Vx3.")
    if faces.dim() != 2 or faces.shape[1] != 3:
        raise ValueError("faces need to be of shape Fx3.")
    return ptk_b200.ops.face_areas_normals(verts, faces)
