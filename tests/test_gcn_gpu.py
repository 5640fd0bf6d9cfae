"""GPU parity of the GCN kernels against the reference's own GCN outputs/gradients (golden) and the
CPU oracle.  Bar: 1e-5 relative, FP32, norm-wise (max-abs error over max-abs reference)."""
import types

import numpy as np
import pytest
import torch

import ptk_b200
from ptk_b200.graph import Graph

pytestmark = pytest.mark.gpu
TOL = 1e-5


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def build_net(g, name, device="cuda"):
    cin, hid, nl, B, N, ignore = [int(v) for v in g[name + "_meta"]]
    args = types.SimpleNamespace(num_GCN_layers=nl, hidden_GCN_size=hid, cut=float(g[name + "_cut"][0]))
    net = ptk_b200.GCN(cin, args, ignore_touch_matrix=bool(ignore))
    sd = {}
    for i in range(nl):
        sd[f"layers.{i}.weight"] = torch.from_numpy(g[f"{name}_w{i}"])
        sd[f"layers.{i}.bias"] = torch.from_numpy(g[f"{name}_b{i}"])
    net.load_state_dict(sd)  # reference state-dict keys/shapes load unchanged
    return net.to(device), nl


@pytest.mark.parametrize("name,tag", [("p_small", "p"), ("g_small", "g"), ("v_orig", "p")])
def test_gcn_vs_reference_module_golden(golden, name, tag):
    g, adj = golden("gcn"), golden("adjacency")
    net, nl = build_net(g, name)
    info = {k: Graph.from_csr(adj[f"{tag}_{k}_rowptr"], adj[f"{tag}_{k}_col"], "cuda").dense() for k in ("origional", "adj")}
    x = torch.from_numpy(g[name + "_x"]).cuda().requires_grad_(True)
    y = net(x, info)
    assert rel_err(y.detach().cpu().numpy(), g[name + "_y"]) < TOL
    (y * torch.from_numpy(g[name + "_gout"]).cuda()).sum().backward()
    assert rel_err(x.grad.cpu().numpy(), g[name + "_gx"]) < TOL
    for i, layer in enumerate(net.layers):
        assert rel_err(layer.weight.grad.cpu().numpy(), g[f"{name}_gw{i}"]) < TOL, i
        assert rel_err(layer.bias.grad.cpu().numpy(), g[f"{name}_gb{i}"]) < TOL, i
    # bias beyond the propagated slice receives exactly zero gradient (vision/model.py:358)
    L = net.layers[0].propagated()
    assert float(net.layers[0].bias.grad[L:].abs().max()) == 0.0


def test_default_20x300_vs_reference_module_golden(golden):
    """Full default GCN (20 layers x 300, N=1949 with hub rows), weights regenerated from the seed."""
    g, adj = golden("gcn"), golden("adjacency")
    torch.manual_seed(1234)
    args = types.SimpleNamespace(num_GCN_layers=20, hidden_GCN_size=300, cut=0.33)
    net = ptk_b200.GCN(50, args)
    x = torch.rand(1, 1949, 50)
    net = net.cuda()
    x = x.cuda().requires_grad_(True)
    info = {"adj": Graph.from_csr(adj["p_adj_rowptr"], adj["p_adj_col"], "cuda").dense()}
    y = net(x, info)
    assert rel_err(y.detach().cpu().numpy(), g["p_default_y"]) < TOL
    gout = torch.rand(1, 1949, 3, generator=torch.Generator().manual_seed(7)).cuda()
    (y * gout).sum().backward()
    assert rel_err(x.grad.cpu().numpy()[:, ::16], g["p_default_gx"]) < TOL
    assert rel_err(net.layers[0].weight.grad.cpu().numpy(), g["p_default_gw0"]) < TOL
    assert rel_err(net.layers[0].bias.grad.cpu().numpy(), g["p_default_gb0"]) < TOL
    assert rel_err(net.layers[19].weight.grad.cpu().numpy(), g["p_default_gw19"]) < TOL


@pytest.mark.parametrize("C,L,relu", [(300, 99, True), (300, 300, False), (100, 33, True), (3, 3, False),
                                       (50, 50, False), (7, 2, True), (64, 0, True)])
def test_aggregate_vs_oracle(oracle, golden, C, L, relu):
    adj = golden("adjacency")
    rp, col = adj["g_adj_rowptr"], adj["g_adj_col"]
    gr = Graph.from_csr(rp, col, "cuda")
    rng = np.random.default_rng(C * 31 + L)
    H = rng.standard_normal((2, gr.n, C)).astype(np.float32)
    bias = rng.standard_normal(C).astype(np.float32)
    out = ptk_b200.ops._aggregate(gr, torch.from_numpy(H).cuda(), L, torch.from_numpy(bias).cuda(), relu)
    want = oracle.gcn_aggregate_fwd(rp, col, H, L, bias, relu)
    assert rel_err(out.cpu().numpy(), want) < TOL
    # transpose gather (backward) vs oracle
    gH = ptk_b200.ops._aggregate(gr, torch.from_numpy(H).cuda(), L, None, False, transpose=True)
    want_g, want_b = oracle.gcn_aggregate_bwd(rp, col, H, L)
    assert rel_err(gH.cpu().numpy(), want_g) < TOL
    gb = ptk_b200.ops._bias_grad(torch.from_numpy(H).cuda().reshape(-1, C), L)
    assert rel_err(gb.cpu().numpy(), want_b) < 2e-5 or L == 0


@pytest.mark.parametrize("M,K,N", [(1, 1, 1), (129, 17, 5), (1949, 50, 300), (3898, 300, 300), (1000, 448, 300),
                                   (2324, 300, 3), (257, 63, 130)])
def test_linear_vs_oracle(oracle, M, K, N):
    rng = np.random.default_rng(M + K + N)
    X = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((K, N)) * 0.1).astype(np.float32)
    G = rng.standard_normal((M, N)).astype(np.float32)
    Xc, Wc, Gc = (torch.from_numpy(a).cuda() for a in (X, W, G))
    H = ptk_b200.ops._linear_fwd(Xc, Wc)
    assert rel_err(H.cpu().numpy(), oracle.gcn_linear(X, W)) < TOL
    gX = ptk_b200.ops._linear_dgrad(Gc, Wc, None)
    assert rel_err(gX.cpu().numpy(), G.astype(np.float64) @ W.astype(np.float64).T) < TOL
    gXm = ptk_b200.ops._linear_dgrad(Gc, Wc, Xc)
    assert rel_err(gXm.cpu().numpy(), (G.astype(np.float64) @ W.astype(np.float64).T) * (X > 0)) < TOL
    gW = ptk_b200.ops._linear_wgrad(Xc, Gc)
    assert rel_err(gW.cpu().numpy(), X.astype(np.float64).T @ G.astype(np.float64)) < TOL


def test_layer_signature_with_lambda_activation(golden, oracle):
    """GCN_layer.forward(features, adj, activation) with a non-ReLU callable (vision/model.py:324)."""
    adj = golden("adjacency")
    gr = Graph.from_csr(adj["v_adj_rowptr"], adj["v_adj_col"], "cuda")
    dense = gr.dense()
    torch.manual_seed(0)
    layer = ptk_b200.GCN_layer(20, 30, cut=0.33, do_cut=True).cuda()
    x = torch.rand(2, gr.n, 20, device="cuda")
    y = layer(x, dense, lambda t: t * 2.0)
    want = oracle.gcn_layer_fwd(x.cpu().numpy(), layer.weight.detach().cpu().numpy(), layer.bias.detach().cpu().numpy(),
                                adj["v_adj_rowptr"], adj["v_adj_col"], 0.33, True, False) * 2.0
    assert rel_err(y.detach().cpu().numpy(), want) < TOL
    y2 = layer(x, dense, torch.nn.functional.relu)
    assert float(y2.min()) >= 0.0


def test_gcn_batch_independence(golden):
    adj = golden("adjacency")
    info = {"adj": Graph.from_csr(adj["p_adj_rowptr"], adj["p_adj_col"], "cuda").dense()}
    torch.manual_seed(2)
    args = types.SimpleNamespace(num_GCN_layers=4, hidden_GCN_size=100, cut=0.33)
    net = ptk_b200.GCN(50, args).cuda()
    x = torch.rand(6, 1949, 50, device="cuda")
    saved = ptk_b200.ops.algo["fwd_infer"]
    try:
        ptk_b200.ops.algo["fwd_infer"] = ptk_b200.ops.GEMM_FFMA   # k-sequential FMA chain per output: bit-identical
        with torch.no_grad():
            y = net(x, info)
            y2 = net(x[2:4].contiguous(), info)
        assert torch.equal(y[2:4], y2)
        # tensor-core inference forward: the (tile, k-block) work split depends on the row count, so partial sums are
        # combined at different points -- equal to rounding, not bitwise
        ptk_b200.ops.algo["fwd_infer"] = ptk_b200.ops.GEMM_AUTO
        with torch.no_grad():
            z = net(x, info)
            z2 = net(x[2:4].contiguous(), info)
        assert rel_err(z[2:4].cpu().numpy(), z2.cpu().numpy()) < 2e-6
        assert rel_err(z.cpu().numpy(), y.cpu().numpy()) < 2e-6
    finally:
        ptk_b200.ops.algo["fwd_infer"] = saved


def test_full_size_config3_properties(golden):
    """BASELINE configs[2] size (B=16, N=1949, 20 x 300 GCN): properties that need no oracle run.
    (1) batch independence: every batch element equals the same element run alone (bit-identical forward);
    (2) gradient additivity: parameter gradients of the batch equal the sum of per-element gradients;
    (3) the factored hub rows (graph.factor_hubs) and the plain CSR give the same aggregate."""
    adj = golden("adjacency")
    g = Graph.from_csr(adj["p_adj_rowptr"], adj["p_adj_col"], "cuda")
    info = {"adj": g.dense()}
    args = types.SimpleNamespace(num_GCN_layers=20, hidden_GCN_size=300, cut=0.33)
    torch.manual_seed(7)
    net = ptk_b200.GCN(448, args).cuda()
    B = 16
    x = torch.rand(B, g.n, 448, device="cuda")
    gout = torch.rand(B, g.n, 3, device="cuda")
    y = net(x, info)
    (y * gout).sum().backward()
    gw_batch = [p.grad.clone() for p in net.parameters()]
    sums = [torch.zeros_like(p) for p in net.parameters()]
    for b in (0, 7, 15):
        yb = net(x[b:b + 1].contiguous(), info)
        assert torch.equal(yb[0], y[b]), b
    for b in range(B):
        net.zero_grad(set_to_none=True)
        yb = net(x[b:b + 1].contiguous(), info)
        (yb * gout[b:b + 1]).sum().backward()
        for s, p in zip(sums, net.parameters()):
            s += p.grad
    for a, s in zip(gw_batch, sums):
        assert rel_err(a.cpu().numpy(), s.cpu().numpy()) < TOL
    # (3) factored vs plain hub handling on a batch that uses the tile kernel
    H = torch.randn(B, g.n, 300, device="cuda")
    bias = torch.randn(300, device="cuda")
    fact = ptk_b200.ops._aggregate(g, H, 99, bias, True)
    assert g.fwd_k.n_common > 1000
    plain = torch.empty_like(H)
    import ctypes as C
    from ptk_b200 import _lib
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    _lib.check(_lib.lib().ptk_gcn_aggregate(p(g.rowptr), p(g.col), p(g.val), p(g.hubs), g.n_hubs, g.n, p(H), B, 300, 99,
                                            p(bias), 1, p(plain), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
               "ptk_gcn_aggregate")
    assert rel_err(fact.cpu().numpy(), plain.cpu().numpy()) < 2e-6
    # (4) linearity of the aggregation in its input
    H2 = torch.randn_like(H)
    lhs = ptk_b200.ops._aggregate(g, H + 2.0 * H2, 99, None, False)
    rhs = ptk_b200.ops._aggregate(g, H, 99, None, False) + 2.0 * ptk_b200.ops._aggregate(g, H2, 99, None, False)
    assert rel_err(lhs.cpu().numpy(), rhs.cpu().numpy()) < TOL


@pytest.mark.parametrize("name", ["p_adj", "g_adj", "p_origional"])
@pytest.mark.parametrize("B,C,L,relu", [(3, 300, 99, True), (9, 100, 99, True), (2, 300, 300, False), (17, 132, 33, True),
                                        (1, 448, 448, False), (5, 300, 3, False), (70, 100, 99, True)])
def test_aggregate_forms_are_bit_identical(golden, name, B, C, L, relu):
    """ptk_gcn_aggregate_tiled: the dense-tile product (every union row of a tile read once, 8 accumulator rows in
    registers) and the shared-memory ring (rows staged by cp.async behind mbarriers) against the L2 gather -- same
    neighbour order, same FMA chain => identical bits; forward and transposed graph, batch groups that do not divide B,
    wide layers walked in column chunks; "auto" must equal them too."""
    adj = golden("adjacency")
    gr = Graph.from_csr(adj[name + "_rowptr"], adj[name + "_col"], "cuda")
    assert gr.fwd_k.tile_uptr is not None and 0 < gr.fwd_k.max_union <= 256
    gen = torch.Generator(device="cuda").manual_seed(B * 1000 + C + L)
    H = torch.randn(B, gr.n, C, device="cuda", generator=gen)
    bias = torch.randn(C, device="cuda", generator=gen)
    res = {}
    try:
        for form in ("l2", "dense", "ring", "auto"):
            ptk_b200.ops.aggregate_form = form
            before = ptk_b200._lib.launch_count()
            a = ptk_b200.ops._aggregate(gr, H, L, bias, relu)
            b = ptk_b200.ops._aggregate(gr, H, L, None, False, transpose=True)
            assert ptk_b200._lib.launch_count() - before == 2
            res[form] = (a, b)
    finally:
        ptk_b200.ops.aggregate_form = "auto"
    for form in ("dense", "ring", "auto"):
        assert torch.equal(res[form][0], res["l2"][0]), form
        assert torch.equal(res[form][1], res["l2"][1]), form


def test_random_graphs_with_and_without_common_hub_sets(oracle):
    """Generic CSR graphs: random weights (hub rows NOT rank-1 => no factorisation), rows longer than the
    128-entry strip without a hub list, and a synthetic star graph whose hubs do share a common set."""
    rng = np.random.default_rng(5)
    n = 700
    for case in ("random_weights", "star"):
        rows = [np.unique(np.append(rng.integers(0, n, rng.integers(1, 12)), i)) for i in range(n)]
        hubs = [10, 300, 650]
        common = np.unique(rng.integers(0, n, 400))
        for h in hubs:
            rows[h] = np.unique(np.concatenate([rows[h], common]))
        rowptr = np.zeros(n + 1, np.int32)
        rowptr[1:] = np.cumsum([len(r) for r in rows])
        col = np.concatenate(rows).astype(np.int32)
        if case == "random_weights":
            val = rng.random(len(col)).astype(np.float32)
            gr = Graph(rowptr, col, val, n, "cuda")
            assert gr.fwd_k.n_common == 0          # not rank-1 on the common set: plain hub CTAs
        else:
            gr = Graph.from_csr(rowptr, col, "cuda")
            val = gr.host["val"]
            assert gr.fwd_k.n_common >= 300      # (the pattern is not symmetric: its transpose has no hub rows)
        H = rng.standard_normal((3, n, 64)).astype(np.float32)
        bias = rng.standard_normal(64).astype(np.float32)
        for L in (64, 21):
            out = ptk_b200.ops._aggregate(gr, torch.from_numpy(H).cuda(), L, torch.from_numpy(bias).cuda(), True)
            dense = np.zeros((n, n), np.float64)
            dense[np.repeat(np.arange(n), np.diff(rowptr)), col] = val
            want = H.astype(np.float64).copy()
            want[:, :, :L] = np.einsum("ij,bjc->bic", dense, H[:, :, :L].astype(np.float64)) + bias[:L]
            want = np.maximum(want, 0)
            assert rel_err(out.cpu().numpy(), want) < TOL, (case, L)
            gT = ptk_b200.ops._aggregate(gr, torch.from_numpy(H).cuda(), L, None, False, transpose=True)
            wantT = H.astype(np.float64).copy()
            wantT[:, :, :L] = np.einsum("ji,bjc->bic", dense, H[:, :, :L].astype(np.float64))
            assert rel_err(gT.cpu().numpy(), wantT) < TOL, (case, L)


@pytest.mark.parametrize("hidden,cin,tag", [(100, 50, "p"), (128, 448, "g"), (300, 448, "p"), (64, 52, "p")])
def test_fused_layer_forward_is_bit_identical_to_the_unfused_path(golden, hidden, cin, tag):
    """The fused training forward (split-epilogue GEMM + strided aggregate + packed ReLU mask for dgrad) must give
    exactly the outputs and gradients of the plain linear -> aggregate sequence, for every width it accepts."""
    adj = golden("adjacency")
    info = {"adj": Graph.from_csr(adj[f"{tag}_adj_rowptr"], adj[f"{tag}_adj_col"], "cuda").dense()}
    args = types.SimpleNamespace(num_GCN_layers=5, hidden_GCN_size=hidden, cut=0.33)
    torch.manual_seed(hidden)
    net = ptk_b200.GCN(cin, args).cuda()
    n = info["adj"].shape[0]
    x = torch.rand(3, n, cin, device="cuda")
    gout = torch.rand(3, n, 3, device="cuda")
    res = {}
    for fused in (True, False):
        ptk_b200.ops.fuse_layers = fused
        ptk_b200.ops.batch_bias_grad = fused   # likewise: slab + batched bias gradients vs one pair of launches per layer
        try:
            xi = x.clone().requires_grad_(True)
            net.zero_grad(set_to_none=True)
            y = net(xi, info)
            (y * gout).sum().backward()
            res[fused] = [y.detach().clone(), xi.grad.clone()] + [p.grad.clone() for p in net.parameters()]
        finally:
            ptk_b200.ops.fuse_layers = True
            ptk_b200.ops.batch_bias_grad = True
    for a, b in zip(res[True], res[False]):
        assert torch.equal(a, b)


# ------------------------------------------------------------------------------------------------------------------
# Inference (torch.no_grad) forward: ops.algo["fwd_infer"] = GEMM_AUTO routes it through the tcgen05 3xTF32 kernel.
# All of BASELINE config 4 (policy scoring) runs this path, so it is pinned against the same reference goldens.
@pytest.mark.parametrize("name,tag", [("p_small", "p"), ("g_small", "g"), ("v_orig", "p")])
@pytest.mark.parametrize("algo", ["auto", "ffma"])
def test_gcn_no_grad_forward_vs_reference_module_golden(golden, name, tag, algo):
    g, adj = golden("gcn"), golden("adjacency")
    net, _ = build_net(g, name)
    info = {k: Graph.from_csr(adj[f"{tag}_{k}_rowptr"], adj[f"{tag}_{k}_col"], "cuda").dense() for k in ("origional", "adj")}
    x = torch.from_numpy(g[name + "_x"]).cuda()
    saved = ptk_b200.ops.algo["fwd_infer"]
    # "auto" (the default) = tcgen05 3xTF32 for every layer the tensor-core kernel accepts (K >= 32, N >= 16), FFMA else
    ptk_b200.ops.algo["fwd_infer"] = {"auto": ptk_b200.ops.GEMM_AUTO, "ffma": ptk_b200.ops.GEMM_FFMA}[algo]
    try:
        with torch.no_grad():
            y = net(x, info)
    finally:
        ptk_b200.ops.algo["fwd_infer"] = saved
    assert not y.requires_grad
    assert rel_err(y.cpu().numpy(), g[name + "_y"]) < TOL


def _default_net_and_input(adj):
    torch.manual_seed(1234)
    args = types.SimpleNamespace(num_GCN_layers=20, hidden_GCN_size=300, cut=0.33)
    net = ptk_b200.GCN(50, args)
    x = torch.rand(1, 1949, 50)
    info = {"adj": Graph.from_csr(adj["p_adj_rowptr"], adj["p_adj_col"], "cuda").dense()}
    return net, x, info


def test_default_20x300_no_grad_forward_vs_reference_module_golden(golden):
    """The full default network (20 x 300, N=1949 with hub rows) through the tensor-core inference forward: 19 of its
    20 linear layers (50->300, 18 x 300->300) run on tcgen05, the 300->3 output layer on the skinny FFMA kernel."""
    g, adj = golden("gcn"), golden("adjacency")
    net, x, info = _default_net_and_input(adj)
    net, x = net.cuda(), x.cuda()
    saved = ptk_b200.ops.algo["fwd_infer"]
    ptk_b200.ops.algo["fwd_infer"] = ptk_b200.ops.GEMM_AUTO
    try:
        with torch.no_grad():
            y = net(x, info)
        # the tensor-core kernel must really be what ran: forcing it on the 300 -> 300 shape succeeds ...
        h = torch.rand(1949, 300, device="cuda")
        w = torch.rand(300, 300, device="cuda")
        tc = ptk_b200.ops._linear_fwd(h, w, algo_id=ptk_b200.ops.GEMM_TF32X3)
        ff = ptk_b200.ops._linear_fwd(h, w, algo_id=ptk_b200.ops.GEMM_FFMA)
        assert not torch.equal(tc, ff) and rel_err(tc.cpu().numpy(), ff.cpu().numpy()) < 2e-6
        assert torch.equal(ptk_b200.ops._linear_fwd(h, w, algo_id=ptk_b200.ops.GEMM_AUTO), tc)
        with pytest.raises(ValueError):  # ... and is refused (not silently replaced) where it does not apply
            ptk_b200.ops._linear_fwd(h, w[:, :3].contiguous(), algo_id=ptk_b200.ops.GEMM_TF32X3)
        launches0 = ptk_b200._lib.launch_count()
        with torch.no_grad():
            net(x, info)
        assert ptk_b200._lib.launch_count() - launches0 >= 40  # 20 x (linear + aggregate) of OUR kernels ran
    finally:
        ptk_b200.ops.algo["fwd_infer"] = saved
    assert rel_err(y.cpu().numpy(), g["p_default_y"]) < TOL


def _h1_report(golden, fwd):
    """Errors of our run and of the reference's fp32 run (the golden) against the same network in fp64
    (oracle/torch_ref.py, dense adjacency as the reference multiplies it)."""
    from oracle import torch_ref as tr
    g, adj = golden("gcn"), golden("adjacency")
    net, x, info = _default_net_and_input(adj)
    gout = torch.rand(1, 1949, 3, generator=torch.Generator().manual_seed(7))
    x64 = x.double().requires_grad_(True)
    ws = [l.weight.detach().double().requires_grad_(True) for l in net.layers]
    bs = [l.bias.detach().double().requires_grad_(True) for l in net.layers]
    y64 = tr.gcn_dense(x64, ws, bs, info["adj"].cpu().double(), 0.33)
    (y64 * gout.double()).sum().backward()
    truth = {"y": y64.detach().numpy(), "gx": x64.grad.numpy()[:, ::16], "gw0": ws[0].grad.numpy(),
             "gb0": bs[0].grad.numpy(), "gw19": ws[19].grad.numpy()}
    ref32 = {"y": g["p_default_y"], "gx": g["p_default_gx"], "gw0": g["p_default_gw0"], "gb0": g["p_default_gb0"],
             "gw19": g["p_default_gw19"]}
    net = net.cuda()
    xg = x.cuda().requires_grad_(True)
    saved = ptk_b200.ops.algo["fwd_train"]
    ptk_b200.ops.algo["fwd_train"] = ptk_b200.ops.GEMM_FFMA if fwd == "ffma" else ptk_b200.ops.GEMM_AUTO
    try:
        y = net(xg, info)
        (y * gout.cuda()).sum().backward()
    finally:
        ptk_b200.ops.algo["fwd_train"] = saved
    ours = {"y": y.detach().cpu().numpy(), "gx": xg.grad.cpu().numpy()[:, ::16],
            "gw0": net.layers[0].weight.grad.cpu().numpy(), "gb0": net.layers[0].bias.grad.cpu().numpy(),
            "gw19": net.layers[19].weight.grad.cpu().numpy()}
    rep = {k: (rel_err(ours[k], truth[k]), rel_err(ref32[k], truth[k])) for k in truth}
    print(fwd, {k: "%.2e (reference fp32: %.2e)" % v for k, v in rep.items()})
    return rep


@pytest.mark.parametrize("fwd", ["ffma", "tensor_core"])
def test_default_20x300_output_error_against_fp64_is_within_twice_the_references_own(golden, fwd):
    """SURVEY H1's second criterion on the OUTPUTS: the reference's own fp32 run has some error against fp64; ours --
    with the exact-FFMA training forward and with the tensor-core 3xTF32 forward -- must not exceed twice it."""
    e_ours, e_ref = _h1_report(golden, fwd)["y"]
    assert e_ours <= 2.0 * max(e_ref, 1e-6), (fwd, e_ours, e_ref)


def test_default_20x300_gradient_error_against_fp64_is_within_twice_the_references_own(golden):
    """The same criterion on the GRADIENTS, for the default (exact FFMA) training forward."""
    for k, (e_ours, e_ref) in _h1_report(golden, "ffma").items():
        assert e_ours <= 2.0 * max(e_ref, 1e-6), (k, e_ours, e_ref)


def test_tensor_core_training_forward_flips_relus_which_is_why_it_is_not_the_default(golden):
    """Measured fact behind ops.algo['fwd_train'] = GEMM_FFMA.  With the 3xTF32 forward the OUTPUT is as close to fp64 as
    the reference's (test above), but a handful of the 11 M ReLU units of the 20-layer stack sit within rounding of
    zero and switch, and every switched unit moves the input gradient by ~1e-4: on this network gx is 6.6e-4 from
    fp64 while the reference's fp32 run (whose masks happen to equal the fp64 run's) is 4.8e-7 from it.  The exact
    forward reproduces the reference's masks and keeps 1e-5.  The tensor-core forward stays opt-in for training
    (bench.py extra.recon_step.tensor_core_forward states what the exact one costs) and is the inference default."""
    rep = _h1_report(golden, "tensor_core")
    e_ours, e_ref = rep["gx"]
    assert e_ours < 1e-2                      # a few flipped units, not a wrong gradient
    assert rep["gw19"][0] < 1e-2 and rep["gw0"][0] < 1e-2
    if e_ours <= 2.0 * max(e_ref, 1e-6):
        pytest.skip("no ReLU flipped on this build: the tensor-core forward met H1 on the gradients as well")


@pytest.mark.parametrize("hidden,cin,layers,tag", [(300, 448, 20, "p"), (100, 50, 5, "g"), (64, 52, 3, "p"), (40, 16, 1, "p")])
def test_native_stack_runner_is_bit_identical_to_the_per_layer_calls(golden, hidden, cin, layers, tag):
    """ptk_gcn_stack_fwd / ptk_gcn_stack_bwd (one C-ABI call per pass, csrc/gcn_stack.cu) against the same layer loop
    driven from Python through the per-layer entry points: outputs, input gradient and every parameter gradient equal
    bit for bit, in training and in inference (tensor-core forward), with and without the fused layer forward."""
    adj = golden("adjacency")
    info = {"adj": Graph.from_csr(adj[f"{tag}_adj_rowptr"], adj[f"{tag}_adj_col"], "cuda").dense()}
    args = types.SimpleNamespace(num_GCN_layers=layers, hidden_GCN_size=hidden, cut=0.33)
    torch.manual_seed(hidden + layers)
    net = ptk_b200.GCN(cin, args).cuda()
    n = info["adj"].shape[0]
    x = torch.rand(2, n, cin, device="cuda")
    gout = torch.rand(2, n, 3, device="cuda")
    res = {}
    try:
        for native in (True, False):
            for fused in (True, False):
                ptk_b200.ops.native_stack, ptk_b200.ops.fuse_layers = native, fused
                xi = x.clone().requires_grad_(True)
                net.zero_grad(set_to_none=True)
                n0 = ptk_b200._lib.launch_count()
                y = net(xi, info)
                (y * gout).sum().backward()
                launches = ptk_b200._lib.launch_count() - n0
                with torch.no_grad():
                    yi = net(x, info)
                res[(native, fused)] = ([y.detach().clone(), xi.grad.clone(), yi.clone()] +
                                        [p.grad.clone() for p in net.parameters()], launches)
    finally:
        ptk_b200.ops.native_stack, ptk_b200.ops.fuse_layers = True, True
    for fused in (True, False):
        (a, la), (b, lb) = res[(True, fused)], res[(False, fused)]
        # same kernels, minus the per-layer weight splits and split-M reductions it batches into one launch per pass
        assert la <= lb and la >= lb - 3 * layers, (la, lb)
        for u, v in zip(a, b):
            assert torch.equal(u, v)


@pytest.mark.parametrize("mode", ["1", "3"])
def test_cluster_forms_of_the_tensor_core_gemm(mode):
    """The 2-CTA cluster forms of the tcgen05 forward / dgrad GEMM (PTK_TG_PAIR=1: weight tiles multicast over the pair,
    3: cta_group::2 MMAs over a 256-row tile; both off by default) give the single-CTA kernel's accuracy
    against fp64 on ragged and full-size shapes.  The switch is read once per process: run in a child interpreter."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PTK_TG_PAIR=mode)
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "tc_2sm_check.py")], env=env, capture_output=True,
                         text=True, timeout=240)
    assert out.returncode == 0, out.stdout + out.stderr


def test_forward_gemm_tail_tiles_are_bit_identical():
    """The exact forward GEMM cuts the rows behind the last full round of CTAs into lower tiles ('tail
    tiles'); PTK_FWD_TAIL=0 keeps 64-row tiles everywhere.  Same per-element arithmetic: the bit-level checksums of
    tools/fwd_exact_check.py (reconstruction-step shapes, one-tile and ragged shapes) must not change.  The switch is read
    once per process: two child interpreters."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for tail in ("1", "0"):
        out = subprocess.run([sys.executable, os.path.join(root, "tools", "fwd_exact_check.py")],
                             env=dict(os.environ, PTK_FWD_TAIL=tail), capture_output=True, text=True, timeout=240)
        assert out.returncode == 0, out.stdout + out.stderr
        outs.append([l for l in out.stdout.splitlines() if "checksum" in l])
    assert len(outs[0]) >= 6 and outs[0] == outs[1]
