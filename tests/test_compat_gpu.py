"""GPU tests of the reference-signature layer (utils.chamfer_distance etc.) and the host-buffer ABI."""
import types

import numpy as np
import pytest
import torch

import ptk_b200

pytestmark = pytest.mark.gpu
TOL = 1e-5


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def test_mesh_chamfer_golden(golden):
    """utils.chamfer_distance(verts, faces, gt, num, repeat=3) vs the reference's own function."""
    g = golden("sampler")
    verts = torch.from_numpy(g["meshcd_verts"]).cuda().requires_grad_(True)
    faces = torch.from_numpy(g["meshcd_faces"].astype(np.int64)).cuda()
    gt = torch.from_numpy(g["meshcd_gt"]).cuda()
    uni = [(torch.from_numpy(g[f"meshcd_u_face{r}"]).cuda(), torch.from_numpy(g[f"meshcd_uv{r}"]).cuda())
           for r in range(3)]
    cd = ptk_b200.utils.chamfer_distance(verts, faces, gt, num=1500, repeat=3, uniforms=uni)
    assert cd.shape == (2,)
    assert rel_err(cd.detach().cpu().numpy(), g["meshcd_cd"]) < TOL
    cd.sum().backward()
    assert rel_err(verts.grad.cpu().numpy(), g["meshcd_gverts"]) < TOL


def test_host_abi_mesh_chamfer_golden(golden):
    g = golden("sampler")
    ctx = ptk_b200.host.HostContext(0)
    u_face = np.stack([g[f"meshcd_u_face{r}"] for r in range(3)])
    uv = np.stack([g[f"meshcd_uv{r}"] for r in range(3)])
    cd, gv = ctx.mesh_chamfer(g["meshcd_verts"], g["meshcd_faces"], g["meshcd_gt"], u_face, uv,
                              grad_cd=np.ones(2, np.float32))
    assert rel_err(cd, g["meshcd_cd"]) < TOL and rel_err(gv, g["meshcd_gverts"]) < TOL
    ctx.close()


def test_host_abi_chamfer(golden, oracle):
    g = golden("chamfer")
    ctx = ptk_b200.host.HostContext(0)
    out = ctx.chamfer(g["ne_x"], g["ne_y"], grad_cham=g["ne_gcham"], want_idx=True)
    ocham, _, oix, _, oiy = oracle.chamfer_fwd(g["ne_x"], g["ne_y"], use_fma=True)
    assert np.array_equal(out["idx_x"], oix) and np.array_equal(out["idx_y"], oiy)
    assert rel_err(out["cham"], g["ne_cham"]) < TOL
    assert rel_err(out["grad_x"], g["ne_gx"]) < TOL and rel_err(out["grad_y"], g["ne_gy"]) < TOL
    with pytest.raises(ValueError):
        ctx.chamfer(np.zeros((1, 0, 3), np.float32), np.zeros((1, 4, 3), np.float32))
    ctx.close()


def test_host_abi_chamfer_chunked_pipeline_matches_device_api():
    """Enough pairs that ptk_host_chamfer splits the batch into upload/compute chunks: the result must be
    bit-identical to the single-launch device API."""
    B, P1, P2 = 900, 2100, 1900
    rng = np.random.default_rng(4)
    x = rng.random((B, P1, 3), np.float32)
    y = rng.random((B, P2, 3), np.float32)
    gc = rng.random(B).astype(np.float32)
    ctx = ptk_b200.host.HostContext(0)
    out = ctx.chamfer(x, y, grad_cham=gc, want_idx=True)
    ctx.close()
    xt, yt = torch.from_numpy(x).cuda().requires_grad_(True), torch.from_numpy(y).cuda().requires_grad_(True)
    cham, ix, iy = ptk_b200.ops.chamfer(xt, yt)
    (cham * torch.from_numpy(gc).cuda()).sum().backward()
    assert np.array_equal(out["idx_x"], ix.cpu().numpy()) and np.array_equal(out["idx_y"], iy.cpu().numpy())
    assert np.array_equal(out["cham"], cham.detach().cpu().numpy())
    assert rel_err(out["grad_x"], xt.grad.cpu().numpy()) < 1e-6 and rel_err(out["grad_y"], yt.grad.cpu().numpy()) < 1e-6


def test_adj_init_on_gpu_and_deformation_like_step(golden, objects_dir):
    """load_mesh_vision -> adj_init -> GCN -> vertex update -> chamfer loss -> backward: the
    sequence of vision/train.py:120-157 with the CNN encoders replaced by random features."""
    args = types.SimpleNamespace(use_touch=True, finger=True, num_grasps=5, num_GCN_layers=3, hidden_GCN_size=60,
                                 cut=0.33)
    adj_info, verts = ptk_b200.utils.load_mesh_vision(args, objects_dir + "/vision_charts.obj")
    assert adj_info["adj"].shape == (1949, 1949) and adj_info["faces"].shape == (2464, 3)
    torch.manual_seed(0)
    net = ptk_b200.GCN(50, args).cuda()
    B = 2
    touch = torch.rand(B, 125, 3, device="cuda") * 0.1
    feats = torch.rand(B, 1949, 50, device="cuda")
    vertices = torch.cat([verts[None].repeat(B, 1, 1), touch], dim=1)
    update = net(feats, adj_info)
    vertices = torch.cat([vertices[:, :1824] + update[:, :1824], vertices[:, 1824:]], dim=1)
    gt = torch.rand(B, 3000, 3, device="cuda") * 0.4 - 0.2
    loss = 9000 * ptk_b200.utils.chamfer_distance(vertices, adj_info["faces"], gt, num=3000).mean()
    loss.backward()
    assert torch.isfinite(loss)
    for layer in net.layers:
        assert torch.isfinite(layer.weight.grad).all() and float(layer.weight.grad.abs().max()) > 0


@pytest.mark.parametrize("algo,B,P1,P2", [("filter", 4, 2000, 1500), ("pruned", 4, 2000, 1500), ("pruned", 1, 20000, 17000),
                                          ("pruned", 70, 33000, 3000)])
def test_cuda_graph_capture_of_forward(algo, B, P1, P2):
    """Every ABI call is capture-safe (no sync, no allocation inside the library) -- the brute-force scan and all three
    sort variants of the pruned scan (one CTA per cloud with 16^3 / 32^3 cells, the multi-launch form for few large clouds)."""
    x = torch.rand(B, P1, 3, device="cuda")
    y = torch.rand(B, P2, 3, device="cuda")
    try:
        ptk_b200.ops.set_chamfer_algo(algo)
        ref, rix, riy = ptk_b200.ops.chamfer(x, y)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            ptk_b200.ops.chamfer(x, y)  # warm-up on the side stream
            with torch.cuda.graph(g, stream=s):
                cham, ix, iy = ptk_b200.ops.chamfer(x, y)
        for _ in range(2):
            g.replay()
        torch.cuda.synchronize()
    finally:
        ptk_b200.ops.set_chamfer_algo("auto")
    assert torch.equal(cham, ref) and torch.equal(ix, rix) and torch.equal(iy, riy)


def test_graphed_step_replays_the_eager_step(golden):
    """recon.GraphedStep: forward + loss + backward + (capturable) Adam of a small chart deformer captured in one CUDA
    graph; with the learning rate held at 0 both arms see the same weights, so a replay's loss and parameter gradients
    must equal the eager step's (atomics in the Chamfer backward: 1e-5).  tools/graph_ddp_check.py is the same check
    at world 2 with the NCCL gradient all-reduce inside the capture."""
    import copy
    from ptk_b200.graph import Graph
    adj, meshes = golden("adjacency"), golden("meshes")
    g = Graph.from_csr(adj["p_adj_rowptr"], adj["p_adj_col"], "cuda")
    adj_info = {"origional": Graph.from_csr(adj["p_origional_rowptr"], adj["p_origional_col"], "cuda").dense(),
                "adj": g.dense(), "faces": torch.from_numpy(adj["p_faces"]).cuda().long()}
    args = types.SimpleNamespace(use_img=True, use_touch=True, finger=True, num_grasps=5, num_GCN_layers=5,
                                 hidden_GCN_size=96, cut=0.33)
    B, width, npts = 3, 40, 2000
    torch.manual_seed(0)
    net_e = ptk_b200.recon.ChartDeformer(adj_info, args, width).cuda()
    net_g = copy.deepcopy(net_e)
    gen = torch.Generator(device="cuda").manual_seed(5)
    vision = torch.from_numpy(meshes["vision_verts"]).cuda()[None].repeat(B, 1, 1)
    touch = torch.rand(B, 125, 3, device="cuda", generator=gen) * 0.02 + 0.2
    feats = [torch.rand(B, n, width, device="cuda", generator=gen) for n in (1824, 1949, 1949)]
    gt = torch.nn.functional.normalize(torch.randn(B, npts, 3, device="cuda", generator=gen), dim=-1) * 0.25
    uni = [ptk_b200.utils.draw_uniforms(B, npts, torch.device("cuda"), gen) for _ in range(3)]
    lr = torch.zeros((), device="cuda")

    def make(net):
        opt = torch.optim.Adam(net.parameters(), lr=lr, fused=True, capturable=True)

        def step():
            opt.zero_grad(set_to_none=True)
            verts = net(vision, touch, lambda it, v: feats[it])
            loss, _ = ptk_b200.recon.recon_loss(verts, adj_info["faces"], gt, number_points=npts, uniforms=uni)
            loss.backward()
            opt.step()
            return loss
        return step

    le = make(net_e)()
    graphed = ptk_b200.recon.GraphedStep(make(net_g), warmup=2)
    for _ in range(2):
        lg = graphed()
    torch.cuda.synchronize()
    assert abs(float(lg.detach()) - float(le.detach())) <= 1e-6 * abs(float(le.detach()))
    for a, b in zip(net_g.parameters(), net_e.parameters()):
        assert torch.equal(a, b)                 # lr = 0: untouched
        assert float((a.grad - b.grad).abs().max()) <= 1e-5 * float(b.grad.abs().max()) + 1e-12


def test_config2_touch_chart_loss_full_size(oracle, golden):
    """BASELINE config 2 shape (touch/train.py:226-240): 64 touch charts (25 vertices, 32 faces) moved by random
    rigid frames, 4000 sampled points vs 4000 ground-truth points, repeat = 3, loss = 9000 * mean.  Loss, per-chart
    distances and the gradient on the chart vertices against the C oracle chained the same way."""
    m = golden("meshes")
    B, S, P = 64, 4000, 4000
    rng = np.random.default_rng(2)
    chart = m["touch_verts"].astype(np.float32)                       # (25, 3)
    faces = m["touch_faces"].astype(np.int64)                         # (32, 3)
    rot = np.linalg.qr(rng.standard_normal((B, 3, 3)))[0].astype(np.float32)
    pos = (0.1 * rng.standard_normal((B, 1, 3))).astype(np.float32)
    verts = (chart[None] @ rot + pos).astype(np.float32)
    verts += (0.0005 * rng.standard_normal(verts.shape)).astype(np.float32)
    u0, uv0 = rng.random((B, P), dtype=np.float32), rng.random((2, B, P), dtype=np.float32)
    gt, _ = oracle.sample_fwd((chart[None] @ rot + pos).astype(np.float32), faces, u0, uv0)
    gt = (gt + 0.002 * rng.standard_normal(gt.shape)).astype(np.float32)
    uni = [(rng.random((B, S), dtype=np.float32), rng.random((2, B, S), dtype=np.float32)) for _ in range(3)]

    vt = torch.from_numpy(verts).cuda().requires_grad_(True)
    cd = ptk_b200.utils.chamfer_distance(vt, torch.from_numpy(faces).cuda(), torch.from_numpy(gt).cuda(), num=S, repeat=3,
                                         uniforms=[(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()) for a, b in uni])
    loss = 9000.0 * cd.mean()
    loss.backward()

    want_cd = np.zeros(B, np.float64)
    want_g = np.zeros_like(verts, dtype=np.float64)
    gcham = np.full(B, 9000.0 / (3 * B), np.float32)
    for u_face, uv in uni:
        pts, fidx = oracle.sample_fwd(verts, faces, u_face, uv)
        cham, _, ix, _, iy = oracle.chamfer_fwd(pts, gt)
        gx, _ = oracle.chamfer_bwd(pts, gt, ix, iy, gcham, want_y=False)
        want_g += oracle.sample_bwd(gx, fidx, uv, faces, verts.shape[1])
        want_cd += cham
    want_cd /= 3
    assert rel_err(cd.detach().cpu().numpy(), want_cd) < TOL
    assert abs(float(loss) - 9000.0 * want_cd.mean()) < TOL * abs(9000.0 * want_cd.mean())
    assert rel_err(vt.grad.cpu().numpy(), want_g) < TOL


@pytest.mark.parametrize("mode", ["uniform", "multinomial"])
def test_fused_mesh_chamfer_equals_the_per_repeat_composition(golden, mode):
    """utils.chamfer_distance as one autograd node (ptk_mesh_chamfer_fwd / bwd) against the same function composed of
    batch_sample + chamfer per repeat + stack + mean (utils.py:204-217): same RNG consumption (same seed => same draws),
    values, vertex gradient and -- the autoencoder's case (autoencoder/train.py:145-150) -- the gradient of the second cloud."""
    m = golden("meshes")
    verts0 = torch.from_numpy(m["obj0_verts"]).cuda()
    faces = torch.from_numpy(m["obj0_faces"].astype(np.int64)).cuda()
    B, num, P2 = 3, 1700, 2100
    scale = 1.0 + 0.1 * torch.arange(B, device="cuda", dtype=torch.float32)[:, None, None]
    w = torch.tensor([1.0, -2.0, 0.5], device="cuda")
    res = []
    try:
        for fused in (True, False):
            ptk_b200.utils.fused_mesh_chamfer = fused
            g = torch.Generator(device="cuda").manual_seed(11)
            verts = (verts0[None] * scale).clone().requires_grad_(True)
            gt = (torch.rand(B, P2, 3, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5)) * 0.2 - 0.1).requires_grad_(True)
            before = ptk_b200._lib.launch_count()
            cd = ptk_b200.utils.chamfer_distance(verts, faces, gt, num=num, repeat=3, generator=g, face_draw=mode)
            (cd * w).sum().backward()
            res.append((cd.detach(), verts.grad.clone(), gt.grad.clone(), ptk_b200._lib.launch_count() - before,
                        torch.rand(4, device="cuda", generator=g)))
    finally:
        ptk_b200.utils.fused_mesh_chamfer = True
    (c1, gv1, gg1, _, tail1), (c2, gv2, gg2, _, tail2) = res
    assert torch.equal(tail1, tail2)                                   # the generator was advanced identically
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    assert rel(c1, c2) < 1e-6 and rel(gv1, gv2) < 1e-6 and rel(gg1, gg2) < 1e-6
    # only the second cloud needs a gradient (the sampled cloud is detached in the autoencoder's loss)
    g = torch.Generator(device="cuda").manual_seed(11)
    gt = gg1.detach().clone().requires_grad_(True)
    cd = ptk_b200.utils.chamfer_distance((verts0[None] * scale).detach(), faces, gt, num=num, repeat=2, generator=g, face_draw=mode)
    cd.sum().backward()
    assert gt.grad is not None and torch.isfinite(gt.grad).all()
