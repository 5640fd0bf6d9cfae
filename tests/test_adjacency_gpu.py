"""Adjacency built on the device (ptk_adj_count / ptk_adj_emit, Graph.from_faces, utils.adj_init on CUDA inputs)
against (1) the CSR of the reference's own utils.adj_init output (tests/golden/adjacency.npz) and (2) the oracle's
dense restatement (oracle/torch_ref.py) on random meshes with twin vertices.  Integer work: everything bit-exact."""
import types

import numpy as np
import pytest
import torch

import ptk_b200
from ptk_b200.graph import Graph, graph_of

pytestmark = pytest.mark.gpu


def _same_graph(g, rowptr, col):
    ref = Graph.from_csr(rowptr, col, "cpu")  # host transposition by lexsort: independent of the builder's shortcut
    for k in ("rowptr", "col", "val", "rowptr_t", "col_t", "val_t"):
        assert np.array_equal(g.host[k], ref.host[k]), k
    assert g.n_hubs == ref.n_hubs and g.n_hubs_t == ref.n_hubs_t
    assert g.fwd_k.n_common == ref.fwd_k.n_common and g.bwd_k.n_common == ref.bwd_k.n_common


@pytest.mark.parametrize("tag,use_touch,finger", [("v", False, False), ("p", True, True), ("g", True, False)])
def test_device_adj_init_matches_reference(golden, objects_dir, tag, use_touch, finger):
    adj = golden("adjacency")
    args = types.SimpleNamespace(use_touch=use_touch, finger=finger, num_grasps=5)
    info, verts = ptk_b200.utils.load_mesh_vision(args, objects_dir + "/vision_charts.obj")
    assert verts.is_cuda and info["adj"].is_cuda
    for which in ("origional", "adj"):
        g = graph_of(info[which])  # registered by adj_init: no dense scan
        _same_graph(g, adj[f"{tag}_{which}_rowptr"], adj[f"{tag}_{which}_col"])
        # the dense tensor handed to the reference's callers is the row-normalised matrix itself
        want = Graph.from_csr(adj[f"{tag}_{which}_rowptr"], adj[f"{tag}_{which}_col"], "cpu").dense()
        assert torch.equal(info[which].cpu(), want)
        # and rebuilding the graph by scanning it gives the same CSR
        assert np.array_equal(Graph.from_dense(info[which]).host["col"], g.host["col"])
    assert np.array_equal(info["faces"].cpu().numpy(), adj[f"{tag}_faces"])
    assert info["faces"].dtype == torch.int64


@pytest.mark.parametrize("seed,n0,F,finger,grasps", [(0, 97, 300, True, 2), (1, 257, 700, False, 1), (2, 40, 60, True, 5),
                                                      (3, 1500, 2900, True, 3)])
def test_random_mesh_with_twins_vs_oracle(golden, objects_dir, seed, n0, F, finger, grasps):
    from oracle import torch_ref as tr
    rng = np.random.default_rng(seed)
    faces = rng.integers(0, n0, size=(F, 3)).astype(np.int64)  # repeated and degenerate faces included
    faces[0] = [n0 - 1, 0, 1]  # calc_adj sizes the graph by faces.max()
    pool = rng.standard_normal((max(n0 // 2, 3), 3)).astype(np.float32)
    pool[0] = [0.0, 0.0, 0.0]
    pool[1] = [-0.0, 0.0, 0.0]  # differs from pool[0] in its bytes only: NOT a twin (tobytes(), utils.py:81)
    pool[2] = [np.nan, 1.0, 2.0]  # equal bytes => twins although NaN != NaN
    verts = pool[rng.integers(0, len(pool), size=n0)]
    verts[rng.integers(0, n0, size=n0 // 4)] = rng.standard_normal((n0 // 4, 3)).astype(np.float32)  # some singles
    m = golden("meshes")
    sheet_v, sheet_f = torch.from_numpy(m["touch_verts"]), torch.from_numpy(m["touch_faces"].astype(np.int64))
    vt, ft = torch.from_numpy(verts), torch.from_numpy(faces)
    dense, faces_ref = tr.adj_fuse_touch(vt, ft, tr.calc_adj(ft), sheet_v, sheet_f, grasps, finger)
    rp, col = tr.dense_to_csr(tr.normalize_adj(dense))
    rp0, col0 = tr.dense_to_csr(tr.normalize_adj(tr.calc_adj(ft)))

    args = types.SimpleNamespace(use_touch=True, finger=finger, num_grasps=grasps)
    info = ptk_b200.utils.adj_init(vt.cuda(), ft.cuda(), args)
    _same_graph(graph_of(info["adj"]), rp, col)
    _same_graph(graph_of(info["origional"]), rp0, col0)
    assert torch.equal(info["adj"].cpu(), tr.normalize_adj(dense))
    assert torch.equal(info["faces"].cpu(), faces_ref)


def test_from_faces_edge_cases():
    dev = torch.device("cuda")
    one = Graph.from_faces(torch.zeros(1, 3, dtype=torch.int64, device=dev), 1)  # a single vertex: the self loop
    assert list(one.host["rowptr"]) == [0, 1] and list(one.host["col"]) == [0] and one.host["val"][0] == 1.0
    # isolated vertices keep their self loop; n need not be a multiple of 32
    g = Graph.from_faces(torch.tensor([[0, 33, 64]], device=dev), 70)
    deg = np.diff(g.host["rowptr"])
    assert deg.sum() == 70 + 6 and deg[0] == 3 and deg[33] == 3 and deg[64] == 3 and deg[69] == 1
    assert list(g.host["col"][g.host["rowptr"][33]:g.host["rowptr"][34]]) == [0, 33, 64]
    # twins without centres are linked to each other only
    pos = torch.tensor([[1.0, 2, 3], [4, 5, 6], [1, 2, 3], [1, 2, 3]], device=dev)
    t = Graph.from_faces(torch.tensor([[0, 1, 4]], device=dev), 5, positions=pos)
    assert list(t.host["col"][t.host["rowptr"][2]:t.host["rowptr"][3]]) == [0, 2, 3]
    with pytest.raises(ValueError):
        Graph.from_faces(torch.tensor([[0, 1, 5]], device=dev), 5)
    with pytest.raises(ValueError):
        Graph.from_faces(torch.tensor([[0, 1, 2]], device=dev), 70000)
    with pytest.raises(RuntimeError):
        Graph.from_faces(torch.tensor([[0, 1, 2]]), 3)


def test_large_graph_rows_sorted_and_symmetric():
    """n = 20 000 (wider than one 32-word pass per row, several scan chunks per thread): properties only."""
    rng = np.random.default_rng(5)
    n, F = 20000, 60000
    faces = torch.from_numpy(rng.integers(0, n, size=(F, 3))).cuda()
    g = Graph.from_faces(faces, n)
    rp, col = g.host["rowptr"], g.host["col"]
    rows = np.repeat(np.arange(n), np.diff(rp))
    key = rows.astype(np.int64) * n + col
    assert np.all(np.diff(key) > 0)  # ascending columns within ascending rows, no duplicates
    f = faces.cpu().numpy()
    e = np.concatenate([f[:, [0, 1]], f[:, [0, 2]], f[:, [1, 2]]])
    want = np.unique(np.concatenate([e[:, 0] * n + e[:, 1], e[:, 1] * n + e[:, 0], np.arange(n) * (n + 1)]))
    assert np.array_equal(key, want)
    deg = np.diff(rp).astype(np.float32)
    assert np.array_equal(g.host["val"], np.repeat(np.float32(1) / deg, np.diff(rp)))
    assert np.array_equal(g.host["val_t"], (np.float32(1) / deg)[col])
