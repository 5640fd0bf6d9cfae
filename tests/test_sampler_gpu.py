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


# ------------------------------------------------------------------------------------------------ A5 stand-alone
def test_face_areas_normals_standalone_vs_oracle_and_formula(oracle, golden):
    """ptk_face_areas_normals (the pytorch3d.ops.mesh_face_areas_normals drop-in, utils.py:164) on its own: areas
    bit-exact against the C oracle, unit normals bit-exact against c / max(|c|, 1e-6) evaluated in fp32 (PyTorch3D's
    face_areas_normals kernel), including zero-area faces, NaN vertices and the PACKED batch form the reference
    calls it with (faces offset by vert_dim * b, utils.py:158-162)."""
    m = golden("meshes")
    rng = np.random.default_rng(3)
    cases = [(m["obj0_verts"][None], m["obj0_faces"]), (m["vision_verts"][None], m["vision_faces"])]
    V, F, B = 500, 1200, 3
    verts = rng.standard_normal((B, V, 3)).astype(np.float32)
    faces = rng.integers(0, V, (F, 3)).astype(np.int64)
    faces[:40, 1] = faces[:40, 0]          # degenerate: two equal corners -> zero area, zero normal
    verts[1, faces[100]] = verts[1, faces[100, 0]]  # a collapsed face in one batch element only
    verts[2, 17] = np.nan
    cases.append((verts, faces))
    for verts, faces in cases:
        Bc, Vc, _ = verts.shape
        packed_v = torch.from_numpy(np.ascontiguousarray(verts.reshape(-1, 3))).cuda()
        packed_f = torch.from_numpy(np.concatenate([faces.astype(np.int64) + Vc * b for b in range(Bc)])).cuda()
        areas, normals = ptk_b200.ops.face_areas_normals(packed_v, packed_f)
        assert areas.shape == (Bc * len(faces),) and normals.shape == (Bc * len(faces), 3)
        want = oracle.face_areas(verts, faces).reshape(-1)
        assert np.array_equal(areas.cpu().numpy(), want, equal_nan=True)
        v = verts.astype(np.float32)
        with np.errstate(invalid="ignore"):
            a = v[:, faces[:, 1]] - v[:, faces[:, 0]]
            b = v[:, faces[:, 2]] - v[:, faces[:, 0]]
            c = np.stack([a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1], a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                          a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]], -1).astype(np.float32)
            n = np.sqrt((c[..., 0] * c[..., 0] + c[..., 1] * c[..., 1]) + c[..., 2] * c[..., 2]).astype(np.float32)
            clamp = np.where(n < np.float32(1e-6), np.float32(1e-6), n)   # NaN stays NaN, as in the kernel
            want_n = (c / clamp[..., None]).astype(np.float32).reshape(-1, 3)
        assert np.array_equal(normals.cpu().numpy(), want_n, equal_nan=True)
        ok = np.isfinite(want_n).all(1) & (n.reshape(-1) > 1e-6)
        assert np.abs(np.linalg.norm(want_n[ok].astype(np.float64), axis=1) - 1).max() < 1e-6
    # the shim function the reference imports checks shapes like PyTorch3D does
    ptk_b200.install_pytorch3d_shim()
    from pytorch3d.ops.mesh_face_areas_normals import mesh_face_areas_normals
    with pytest.raises(ValueError):
        mesh_face_areas_normals(torch.zeros(5, 2, device="cuda"), torch.zeros(1, 3, dtype=torch.int64, device="cuda"))
    a, nrm = mesh_face_areas_normals(packed_v, packed_f)
    assert torch.equal(torch.nan_to_num(a, nan=-1.0), torch.nan_to_num(areas, nan=-1.0))


# ------------------------------------------------------------------------------------------------ SURVEY 9.3
def test_face_pick_semantics_against_live_torch_multinomial():
    """Where the explicit-uniform pick and ATen's `Tensor.multinomial(num, replacement=True)` (utils.py:170) agree and
    where they cannot, checked against LIVE torch on this GPU with crafted distributions:

    (a) zero-probability faces are NEVER drawn by either (ATen backs off from a zero bucket; our integer weights give a
        zero-area face an empty interval);
    (b) an all-zero-area mesh: the reference's NaN guard (utils.py:166-168) turns the weights into all ones and ATen
        draws uniformly; ours draws uniformly too;
    (c) the empirical face frequencies of both follow the areas (chi-square against the exact probabilities);
    (d) the STREAMS are not interchangeable in general: ATen draws inside its kernel (one Philox subsequence per
        launched thread, first of four outputs used, launch geometry = f(distributions, samples)) and searches a
        float prefix sum whose summation order is CUB's.  Measured on this box (torch 2.11): for ONE distribution and a
        sample count that fits one launch wave, `torch.rand(1, S)` after the same seed yields the very uniforms
        multinomial consumed, and the face histograms even coincide; for batches the thread -> (row, sample) maps of
        the two kernels differ.  Hence the two contracts of ptk_b200.utils.batch_sample: explicit uniforms (bit-exact
        by construction) and `face_draw="multinomial"` (the draw itself done by ATen, see
        test_batch_sample_multinomial_mode_reproduces_the_references_stream)."""
    dev = "cuda"
    # disjoint right triangles with prescribed areas: face f = (0,0,f), (1,0,f), (1,w_f,f) -> area w_f / 2
    w = torch.tensor([0.0, 3.0, 0.0, 0.0, 1.0, 5.0, 0.0, 1e-3, 2.0, 0.0], device=dev)
    F = w.numel()
    z = torch.arange(F, device=dev, dtype=torch.float32)
    zero_, one_ = torch.zeros(F, device=dev), torch.ones(F, device=dev)
    verts = torch.stack([torch.stack([zero_, zero_, z], 1), torch.stack([one_, zero_, z], 1),
                         torch.stack([one_, w, z], 1)], 1).reshape(1, 3 * F, 3).contiguous()
    faces = torch.arange(3 * F, device=dev).reshape(F, 3)
    areas, _ = ptk_b200.ops.face_areas_normals(verts[0], faces)
    assert torch.allclose(areas, w / 2, rtol=1e-6, atol=0)
    S = 200000
    gen = torch.Generator(device=dev).manual_seed(5)
    u_face = torch.rand(1, S, device=dev, generator=gen)
    uv = torch.rand(2, 1, S, device=dev, generator=gen)
    _, fidx = ptk_b200.ops.sample_points(verts, faces.to(torch.int32), u_face, uv)
    ours = torch.bincount(fidx[0].long(), minlength=F).double()
    p = (areas / areas.sum()).double()
    torch.manual_seed(5)
    aten = torch.bincount(p.float()[None].multinomial(S, replacement=True)[0], minlength=F).double()
    zero = (w == 0)
    assert ours[zero].sum() == 0 and aten[zero].sum() == 0                                  # (a)
    for counts in (ours, aten):                                                            # (c)
        chi2 = (((counts - S * p) ** 2)[~zero] / (S * p)[~zero]).sum().item()
        assert chi2 < 30.0, chi2            # 4 degrees of freedom: P(chi2 > 30) ~ 5e-6
    print("(d) faces drawn, ours vs ATen (same seed):", ours.tolist(), aten.tolist())
    # exact boundary behaviour of OUR pick: u = 0 -> the first face with non-zero area, u -> 1 -> the last one
    edge = torch.tensor([[0.0, 1.0 - 2.0 ** -24]], device=dev)
    _, fe = ptk_b200.ops.sample_points(verts, faces.to(torch.int32), edge, torch.rand(2, 1, 2, device=dev))
    assert fe[0].tolist() == [1, 8]
    # (b) all faces degenerate: uniform fallback in both
    flat = torch.zeros_like(verts)
    _, fz = ptk_b200.ops.sample_points(flat, faces.to(torch.int32), u_face, uv)
    oz = torch.bincount(fz[0].long(), minlength=F).double()
    ar = torch.zeros(1, F, device=dev)
    ar = (ar / ar.sum(1, keepdim=True)).abs()
    ar[ar != ar] = 1                                      # utils.py:166-168
    az = torch.bincount(ar.multinomial(S, replacement=True)[0], minlength=F).double()
    for counts in (oz, az):
        assert (((counts - S / F) ** 2) / (S / F)).sum().item() < 40.0   # 9 dof
