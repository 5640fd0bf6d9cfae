"""GPU parity of the surface sampler: bit-exact points for a given uniform stream."""
import numpy as np
import pytest
import torch

import ptk_b200

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def gpu_sample(verts, faces, u_face, uv, gpts=None):
    v = torch.from_numpy(verts).cuda().requires_grad_(gpts is not None)
    f = torch.from_numpy(faces.astype(np.int32)).cuda()
    pts, fidx = ptk_b200.ops.sample_points(v, f, torch.from_numpy(u_face).cuda(), torch.from_numpy(uv).cuda())
    out = [pts.detach().cpu().numpy(), fidx.cpu().numpy()]
    if gpts is not None:
        (pts * torch.from_numpy(gpts).cuda()).sum().backward()
        out.append(v.grad.cpu().numpy())
    return out


@pytest.mark.parametrize("tag", ["obj0", "p_mesh", "degenerate", "touch"])
def test_golden_bit_exact(golden, tag):
    g = golden("sampler")
    pts, fidx, gv = gpu_sample(g[tag + "_verts"], g[tag + "_faces"], g[tag + "_u_face"], g[tag + "_uv"],
                               g[tag + "_gpts"])
    assert np.array_equal(pts, g[tag + "_pts"])  # bit-exact vs the reference's batch_sample
    assert rel_err(gv, g[tag + "_gverts"]) < 1e-5


@pytest.mark.parametrize("B,V,F,S", [(1, 3, 1, 1), (2, 50, 77, 1000), (3, 300, 5000, 4097), (16, 1949, 2464, 10000),
                                     (1, 9000, 17000, 2000)])
def test_random_meshes_vs_oracle(oracle, B, V, F, S):
    rng = np.random.default_rng(V * 7 + F)
    verts = rng.standard_normal((B, V, 3)).astype(np.float32)
    faces = rng.integers(0, V, (F, 3)).astype(np.int32)
    if F > 10:
        faces[3] = faces[3, 0]  # a degenerate face
    u_face = rng.random((B, S), np.float32)
    uv = rng.random((2, B, S), np.float32)
    gp = rng.standard_normal((B, S, 3)).astype(np.float32)
    pts, fidx, gv = gpu_sample(verts, faces, u_face, uv, gp)
    opts, ofidx = oracle.sample_fwd(verts, faces, u_face, uv)
    assert np.array_equal(fidx, ofidx) and np.array_equal(pts, opts)
    ogv = oracle.sample_bwd(gp, ofidx, uv, faces, V)
    assert rel_err(gv, ogv) < 1e-5


def test_edge_uniforms(oracle, golden):
    m = golden("meshes")
    verts, faces = m["obj0_verts"][None], m["obj0_faces"]
    u_face = np.array([[0.0, np.nextafter(np.float32(1), np.float32(0)), 0.5, 2.0 ** -24]], np.float32)
    uv = np.array([[[0.0, 1 - 2.0 ** -24, 0.25, 0.0]], [[0.0, 0.0, 1 - 2.0 ** -24, 0.5]]], np.float32)
    pts, fidx = gpu_sample(verts, faces, u_face, uv)
    opts, ofidx = oracle.sample_fwd(verts, faces, u_face, uv)
    assert np.array_equal(fidx, ofidx) and np.array_equal(pts, opts)


def test_nan_and_zero_area_guards(oracle):
    verts = np.zeros((3, 6, 3), np.float32)
    verts[0] = np.random.default_rng(0).random((6, 3))
    verts[0, 5] = np.nan             # faces touching vertex 5 have NaN area -> weight 0
    verts[1] = 0.25                  # all-degenerate mesh -> uniform fallback
    verts[2] = np.random.default_rng(1).random((6, 3))
    faces = np.array([[0, 1, 2], [1, 2, 3], [2, 3, 5], [0, 0, 0], [3, 4, 0]], np.int32)
    rng = np.random.default_rng(2)
    u_face, uv = rng.random((3, 512), np.float32), rng.random((2, 3, 512), np.float32)
    pts, fidx = gpu_sample(verts, faces, u_face, uv)
    opts, ofidx = oracle.sample_fwd(verts, faces, u_face, uv)
    assert np.array_equal(fidx, ofidx)
    assert np.array_equal(pts, opts, equal_nan=True)
    assert not np.isin(fidx[0], [2, 3]).any() and not (fidx[2] == 3).any()
    assert set(np.unique(fidx[1])) == {0, 1, 2, 3, 4}


def test_distribution_is_area_weighted(golden):
    m = golden("meshes")
    verts, faces = m["obj0_verts"][None], m["obj0_faces"]
    g = torch.Generator(device="cuda").manual_seed(0)
    S = 400000
    u_face = torch.rand(1, S, device="cuda", generator=g)
    uv = torch.rand(2, 1, S, device="cuda", generator=g)
    _, fidx = ptk_b200.ops.sample_points(torch.from_numpy(verts).cuda(), torch.from_numpy(faces).cuda(), u_face, uv)
    counts = torch.bincount(fidx[0].long(), minlength=faces.shape[0]).double().cpu().numpy()
    v = verts[0].astype(np.float64)
    a = 0.5 * np.linalg.norm(np.cross(v[faces[:, 1]] - v[faces[:, 0]], v[faces[:, 2]] - v[faces[:, 0]]), axis=1)
    p = a / a.sum()
    z = (counts - S * p) / np.sqrt(S * p * (1 - p) + 1e-12)
    assert np.abs(z).max() < 6.0


def test_batch_sample_signature_and_rng_stream(golden):
    """utils.batch_sample(verts, faces, num): int64 faces, (B,num,3) out, consumes the CUDA generator
    as multinomial-draw then rand(2,bs,num); same seed -> same points."""
    m = golden("meshes")
    verts = torch.from_numpy(m["obj0_verts"]).cuda()[None].repeat(2, 1, 1).requires_grad_(True)
    faces = torch.from_numpy(m["obj0_faces"].astype(np.int64)).cuda()
    torch.manual_seed(5)
    p1 = ptk_b200.utils.batch_sample(verts, faces, num=777)
    torch.manual_seed(5)
    u_face, uv = ptk_b200.utils.draw_uniforms(2, 777, verts.device)
    p2 = ptk_b200.utils.batch_sample(verts, faces, num=777, uniforms=(u_face, uv))
    assert p1.shape == (2, 777, 3) and torch.equal(p1, p2)
    p1.sum().backward()
    assert verts.grad is not None and verts.grad.shape == verts.shape
