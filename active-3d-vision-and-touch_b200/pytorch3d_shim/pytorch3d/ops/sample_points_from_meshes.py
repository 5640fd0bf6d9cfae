"""pytorch3d.ops.sample_points_from_meshes._rand_barycentric_coords (utils.py:22,179)."""
import torch


def _rand_barycentric_coords(size1, size2, dtype, device):
    uv = torch.rand(2, size1, size2, dtype=dtype, device=device)
    u, v = uv[0], uv[1]
    u_sqrt = u.sqrt()
    w0 = 1.0 - u_sqrt
    w1 = u_sqrt * (1.0 - v)
    w2 = u_sqrt * v
    return w0, w1, w2
