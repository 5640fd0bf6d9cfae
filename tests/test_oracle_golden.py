"""Pins the CPU oracle (oracle/ptk_oracle.c) against the golden vectors generated from the reference's
own code (oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import torch_ref as tr


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


# ---------------------------------------------------------------- Chamfer / 1-NN
def test_config1_indices_and_value(oracle, golden):
    g = golden("chamfer")
    cham, dx, ix, dy, iy = oracle.chamfer_fwd(g["c1_x"], g["c1_y"], use_fma=False)
    assert np.array_equal(ix, g["c1_idx_x"]) and np.array_equal(iy, g["c1_idx_y"])
    assert rel_err(cham, g["c1_cham"]) < 1e-5
    # the CUDA-order (FMA) arithmetic: indices identical on config 1, value within tolerance
    cham_f, _, ixf, _, iyf = oracle.chamfer_fwd(g["c1_x"], g["c1_y"], use_fma=True)
    assert np.array_equal(ixf, g["c1_fma_idx_x"]) and np.array_equal(iyf, g["c1_fma_idx_y"])
    assert np.array_equal(ixf, g["c1_idx_x"]) and np.array_equal(iyf, g["c1_idx_y"])
    assert rel_err(cham_f, g["c1_cham"]) < 1e-5


def test_ties_lowest_index_wins(oracle, golden):
    g = golden("chamfer")
    for fma in (False, True):
        cham, _, ix, _, iy = oracle.chamfer_fwd(g["tie_x"], g["tie_y"], use_fma=fma)
        assert np.array_equal(ix, g["tie_idx_x"]) and np.array_equal(iy, g["tie_idx_y"])
        assert ix[0, :200].max() < 300  # the first of the 4 tiled copies
        assert rel_err(cham, g["tie_cham"]) < 1e-5


def test_unequal_sizes_and_gradients(oracle, golden):
    g = golden("chamfer")
    cham, _, ix, _, iy = oracle.chamfer_fwd(g["ne_x"], g["ne_y"], use_fma=True)
    assert rel_err(cham, g["ne_cham"]) < 1e-5
    gx, gy = oracle.chamfer_bwd(g["ne_x"], g["ne_y"], ix, iy, g["ne_gcham"])
    assert rel_err(gx, g["ne_gx"]) < 1e-5 and rel_err(gy, g["ne_gy"]) < 1e-5
    gx2, gy2 = oracle.chamfer_bwd(g["ne_x"], g["ne_y"], ix, iy, g["ne_gcham"], want_x=False)
    assert gx2 is None and np.array_equal(gy2, gy)


def test_single_point_clouds(oracle):
    x = np.array([[[0.0, 0.0, 0.0]]], np.float32)
    y = np.array([[[1.0, 2.0, 2.0]]], np.float32)
    cham, dx, ix, dy, iy = oracle.chamfer_fwd(x, y)
    assert cham[0] == 18.0 and ix[0, 0] == 0 and iy[0, 0] == 0


# ---------------------------------------------------------------- sampler
@pytest.mark.parametrize("tag", ["obj0", "p_mesh", "degenerate", "touch"])
def test_sampler_bit_exact_vs_reference_batch_sample(oracle, golden, tag):
    g = golden("sampler")
    pts, fidx = oracle.sample_fwd(g[tag + "_verts"], g[tag + "_faces"], g[tag + "_u_face"], g[tag + "_uv"])
    assert np.array_equal(pts, g[tag + "_pts"])
    gv = oracle.sample_bwd(g[tag + "_gpts"], fidx, g[tag + "_uv"], g[tag + "_faces"], g[tag + "_verts"].shape[1])
    assert rel_err(gv, g[tag + "_gverts"]) < 1e-5


def test_zero_area_faces_never_sampled(oracle, golden):
    g = golden("sampler")
    verts, faces = g["p_mesh_verts"], g["p_mesh_faces"]
    _, fidx = oracle.sample_fwd(verts, faces, g["p_mesh_u_face"], g["p_mesh_uv"])
    areas = oracle.face_areas(verts, faces)
    assert (areas == 0).any()
    for b in range(verts.shape[0]):
        assert (areas[b, fidx[b]] > 0).all()


def test_degenerate_mesh_is_uniform(oracle, golden):
    g = golden("sampler")
    _, fidx = oracle.sample_fwd(g["degenerate_verts"], g["degenerate_faces"], g["degenerate_u_face"],
                                g["degenerate_uv"])
    F = g["degenerate_faces"].shape[0]
    expect = np.minimum((g["degenerate_u_face"] * F).astype(np.int64), F - 1)
    assert np.array_equal(fidx, expect)


def test_cumweights_numpy_restatement(oracle):
    rng = np.random.default_rng(0)
    areas = rng.random((3, 257)).astype(np.float32)
    areas[0, 5] = np.nan
    areas[1, :] = 0.0
    areas[2, 9] = np.inf
    assert np.array_equal(oracle.face_cumweights(areas), tr.face_cumweights(areas))


def test_mesh_chamfer_repeat_mean(oracle, golden):
    """utils.chamfer_distance (repeat=3) composed from oracle pieces equals the reference run."""
    g = golden("sampler")
    verts, faces, gt = g["meshcd_verts"], g["meshcd_faces"], g["meshcd_gt"]
    cds, gvs = [], []
    for r in range(3):
        pts, fidx = oracle.sample_fwd(verts, faces, g[f"meshcd_u_face{r}"], g[f"meshcd_uv{r}"])
        cham, _, ix, _, iy = oracle.chamfer_fwd(pts, gt, use_fma=True)
        gx, _ = oracle.chamfer_bwd(pts, gt, ix, iy, np.full(2, 1.0 / 3.0, np.float32), want_y=False)
        gvs.append(oracle.sample_bwd(gx, fidx, g[f"meshcd_uv{r}"], faces, verts.shape[1]))
        cds.append(cham)
    assert rel_err(np.mean(cds, 0), g["meshcd_cd"]) < 1e-5
    assert rel_err(np.sum(gvs, 0), g["meshcd_gverts"]) < 1e-5


# ---------------------------------------------------------------- GCN
@pytest.mark.parametrize("name,akey", [("p_small", "p_adj"), ("g_small", "g_adj"), ("v_orig", "p_origional")])
def test_gcn_forward_vs_reference_module(oracle, golden, name, akey):
    g, adj = golden("gcn"), golden("adjacency")
    cin, hid, nl, B, N, _ = g[name + "_meta"]
    cut = float(g[name + "_cut"][0])
    rp, col = adj[akey + "_rowptr"], adj[akey + "_col"]
    x = g[name + "_x"]
    for i in range(nl):
        x = oracle.gcn_layer_fwd(x, g[f"{name}_w{i}"], g[f"{name}_b{i}"], rp, col, cut, i < nl - 1, i < nl - 1)
    assert rel_err(x, g[name + "_y"]) < 1e-5


def test_gcn_aggregate_backward_vs_reference_autograd(oracle, golden):
    """Last layer of p_small: gW, gb from the oracle pieces equal the reference's autograd."""
    g, adj = golden("gcn"), golden("adjacency")
    name = "p_small"
    cin, hid, nl, B, N, _ = g[name + "_meta"]
    cut = float(g[name + "_cut"][0])
    rp, col = adj["p_adj_rowptr"], adj["p_adj_col"]
    acts = [g[name + "_x"]]
    for i in range(nl):
        acts.append(oracle.gcn_layer_fwd(acts[-1], g[f"{name}_w{i}"], g[f"{name}_b{i}"], rp, col, cut,
                                         i < nl - 1, i < nl - 1))
    gout = g[name + "_gout"]
    gH, gb = oracle.gcn_aggregate_bwd(rp, col, gout, 3)
    assert rel_err(gb, g[f"{name}_gb{nl - 1}"]) < 1e-5
    X = acts[nl - 1].reshape(B * N, -1).astype(np.float64)
    gW = X.T @ gH.reshape(B * N, -1).astype(np.float64)
    assert rel_err(gW, g[f"{name}_gw{nl - 1}"][0]) < 1e-5


def test_torch_dense_restatement_matches_reference(golden):
    g, adj = golden("gcn"), golden("adjacency")
    name = "g_small"
    cin, hid, nl, B, N, _ = g[name + "_meta"]
    from ptk_b200.graph import Graph
    dense = Graph.from_csr(adj["g_adj_rowptr"], adj["g_adj_col"], "cpu").dense()
    ws = [torch.from_numpy(g[f"{name}_w{i}"]) for i in range(nl)]
    bs = [torch.from_numpy(g[f"{name}_b{i}"]) for i in range(nl)]
    y = tr.gcn_dense(torch.from_numpy(g[name + "_x"]), ws, bs, dense, float(g[name + "_cut"][0]))
    assert rel_err(y.numpy(), g[name + "_y"]) < 1e-6


@pytest.mark.parametrize("tag", ["small", "c3"])
def test_nerf_embedding_restatement_matches_reference(golden, tag):
    """oracle/torch_ref.nerf_embedding vs the reference's own Positional_Encoder.nerf_embedding output."""
    from oracle import torch_ref as tr
    g = golden("encoder")
    pos = torch.from_numpy(g[f"{tag}_pos"]).reshape(-1, 3)
    emb = tr.nerf_embedding(pos)
    assert emb.shape == (pos.shape[0], 63)
    assert np.array_equal(emb[:, :60].numpy(), g[f"{tag}_embedding"])
    assert torch.equal(emb[:, 60:], pos)
