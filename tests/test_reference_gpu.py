"""The reference's OWN modules on the B200 path (north_star: "existing training, eval and policy
scripts pick up the new path without edits").

An unedited `pterotactyl` (baseline/_ref, written by tools/install_reference.py; or $PTK_REFERENCE) is imported with
the pytorch3d shim + import stubs and run twice on the same GPU:

  reference arm : its own GCN / GCN_layer / Positional_Encoder / Deformation classes, dense adjacency from its own
                  adj_init, torch GEMMs; Chamfer loss from oracle/torch_ref.py (PyTorch3D is not installable)
  B200 arm      : the same scripts after ptk_b200.install() -- nothing else changes

S1  the reference's own `utils.batch_sample` / `utils.chamfer_distance` BODIES over the pytorch3d shim
S2  the reference's own `Deformation` (BASELINE config 3, `v_t_p`: N=1824 'origional' graph first, touch charts
    concatenated afterwards) forward + loss + backward, and its `Engine.train` method for three Adam steps
INTEGRATION.md's snippets are executed verbatim by tests/test_reference_import.py (no GPU needed).

Tolerances: outputs, losses 1e-5 relative (north_star).  Gradients of the 20-layer network: 1e-5 against the fp32
reference arm OR the SURVEY-H1 criterion (error against an fp64 run of the reference <= 2x the reference's own fp32
error): a single ReLU flipping in a 20-layer stack moves gradients by more than 1e-5 in either arm.
"""
import numpy as np
import pytest
import torch

import ptk_b200

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def ref():
    if ptk_b200.find_reference() is None:
        pytest.skip("no pterotactyl checkout: run tools/install_reference.py (needs /root/reference) or set PTK_REFERENCE")
    import ref_harness as H
    H.strict_fp32()
    r = H.Reference()
    yield r
    ptk_b200.uninstall()


# ----------------------------------------------------------------------------------------------- S1
def test_s1_unedited_utils_bodies_over_the_shim(ref):
    """utils.py:152-217 run literally (ATen multinomial and all); only the five pytorch3d names are ours."""
    import ref_harness as H
    from ref_harness import rel_err
    ptk_b200.uninstall()
    U = ref.utils
    assert U.batch_sample.__module__ == "pterotactyl.utility.utils"  # the reference's body, not ours
    assert U.cuda_cd.__module__.startswith("pytorch3d")
    verts, faces = U.load_mesh_touch(ref.chart("test_objects/0.obj"))  # shim load_obj, then .cuda() (utils.py:194-200)
    assert verts.is_cuda and faces.dtype == torch.int64 and verts.shape == (242, 3) and faces.shape == (544, 3)
    B, num = 3, 5000
    scale = torch.tensor([1.0, 1.3, 0.7], device="cuda")[:, None, None]
    v_shim = (verts[None] * scale).clone().requires_grad_(True)
    v_rest = v_shim.detach().clone().requires_grad_(True)
    gt = U.batch_sample(verts[None].repeat(B, 1, 1) * 1.05, faces, num=4000).detach()

    torch.manual_seed(7)
    pts_shim = U.batch_sample(v_shim, faces, num=num)
    torch.manual_seed(11)
    cd_shim = U.chamfer_distance(v_shim, faces, gt, num=num)
    cd_shim.sum().backward()

    with H.restated_pytorch3d(U):  # the same bodies with the PyTorch3D names rebound to the eager-torch restatement
        torch.manual_seed(7)
        pts_rest = U.batch_sample(v_rest, faces, num=num)
        torch.manual_seed(11)
        cd_rest = U.chamfer_distance(v_rest, faces, gt, num=num)
        cd_rest.sum().backward()
    assert torch.equal(pts_shim, pts_rest), "areas from ptk_face_areas_normals changed the multinomial draw"
    assert rel_err(cd_shim, cd_rest) < TOL
    assert rel_err(v_shim.grad, v_rest.grad) < TOL


def test_batch_sample_multinomial_mode_reproduces_the_references_stream(ref):
    """north_star: "sampled points must be bit-exact given the same RNG stream" -- with the reference's REAL stream.
    ptk_b200.utils.batch_sample(face_draw="multinomial") (what install() selects) against the reference's own
    `utils.batch_sample` body (utils.py:152-187: Tensor.multinomial + torch.rand from the global CUDA generator) run
    from the same seed: identical points, bit for bit; utils.chamfer_distance likewise (three draws in a row)."""
    import ref_harness as H
    ptk_b200.uninstall()
    U = ref.utils
    args = H.c3_args()
    info = ptk_b200.utils.load_mesh_vision(args, ref.chart())[0]
    faces = info["faces"]                                  # the fused 2464-face list of config 3
    g = torch.Generator().manual_seed(2)
    B, V = 4, 1949
    verts = (torch.randn(B, V, 3, generator=g) * 0.2).cuda()
    verts[1, 1824:1849] = verts[1, 1824]                   # an empty touch chart: 32 zero-area faces (environment.py:308)
    verts[2] = 0.0                                         # an all-degenerate mesh: the NaN -> 1 uniform fallback
    gt = torch.randn(B, 3000, 3, generator=g).cuda() * 0.2
    for seed, num in ((5, 10000), (6, 1000), (7, 30000)):
        torch.manual_seed(seed)
        with H.restated_pytorch3d(U):
            want = U.batch_sample(verts, faces, num=num)
        state_after = torch.cuda.get_rng_state()
        torch.manual_seed(seed)
        got = ptk_b200.utils.batch_sample(verts, faces, num=num, face_draw="multinomial")
        assert torch.equal(got, want), (seed, num)
        assert torch.equal(torch.cuda.get_rng_state(), state_after)   # and the generator is left where the reference leaves it
    v1 = verts.clone().requires_grad_(True)
    v2 = verts.clone().requires_grad_(True)
    torch.manual_seed(9)
    with H.restated_pytorch3d(U):
        cd_ref = U.chamfer_distance(v1, faces, gt, num=4000)
    cd_ref.sum().backward()
    torch.manual_seed(9)
    cd = ptk_b200.utils.chamfer_distance(v2, faces, gt, num=4000, face_draw="multinomial")
    cd.sum().backward()
    assert H.rel_err(cd, cd_ref) < TOL and H.rel_err(v2.grad, v1.grad) < TOL
    # explicit uniforms override the mode; an unknown mode is refused
    with pytest.raises(ValueError):
        ptk_b200.utils.batch_sample(verts, faces, num=10, face_draw="philox")


# ----------------------------------------------------------------------------------------------- S2
@pytest.fixture(scope="module")
def c3(ref):
    """Both arms (+ an fp64 run of the reference arm) of one config-3 step, built once."""
    import ref_harness as H
    args = H.c3_args()
    batch = H.make_batch(args, B=2, seed=0)
    out = {"args": args, "batch": batch}
    info, mesh, net = ref.build(args, patched=False)
    out["state"] = {k: v.detach().clone() for k, v in net.state_dict().items()}
    out["ref_classes"] = (type(net.mesh_deform_1).__module__, type(net.positional_encoder).__module__)
    out["ref_adj_shapes"] = (tuple(info["origional"].shape), tuple(info["adj"].shape), tuple(info["faces"].shape))
    out["ref"] = ref.step(args, net, info, mesh, batch, patched=False)
    del net
    info64, mesh64, net64 = ref.build(args, patched=False, state=out["state"], dtype=torch.float64)
    out["ref64"] = ref.step(args, net64, info64, mesh64, batch, patched=False)
    del net64
    info, mesh, net = ref.build(args, patched=True, state=out["state"])
    out["b200_classes"] = (type(net.mesh_deform_1).__module__, type(net.positional_encoder).__module__)
    out["b200_adj_shapes"] = (tuple(info["origional"].shape), tuple(info["adj"].shape), tuple(info["faces"].shape))
    out["b200_graphs"] = (ptk_b200.graph.graph_of(info["origional"]).n, ptk_b200.graph.graph_of(info["adj"]).n)
    out["b200"] = ref.step(args, net, info, mesh, batch, patched=True)
    del net
    ptk_b200.uninstall()
    torch.cuda.empty_cache()
    return out


def test_s2_install_swaps_the_classes_the_reference_builds(c3):
    assert c3["ref_classes"] == ("pterotactyl.reconstruction.vision.model",) * 2
    assert c3["b200_classes"] == ("ptk_b200.model", "ptk_b200.encoders")
    # v_t_p: first deformation on the 1824-vertex 'origional' graph, then the fused 1949-vertex one; 2464 faces
    assert c3["ref_adj_shapes"] == ((1824, 1824), (1949, 1949), (2464, 3)) == c3["b200_adj_shapes"]
    assert c3["b200_graphs"] == (1824, 1949)


def test_s2_reference_deformation_v_t_p_forward_and_loss(c3):
    from ref_harness import rel_err
    v_ref, l_ref, _ = c3["ref"]
    v_b, l_b, _ = c3["b200"]
    assert v_b.shape == v_ref.shape == (2, 1949, 3)
    assert rel_err(v_b, v_ref) < TOL
    assert rel_err(l_b, l_ref) < TOL
    # and both sit on the fp64 run of the reference
    assert rel_err(v_b, c3["ref64"][0]) < 2 * max(rel_err(v_ref, c3["ref64"][0]), TOL)


def test_s2_reference_deformation_v_t_p_gradients(c3):
    from ref_harness import grad_report
    g_ref, g_b, g_64 = c3["ref"][2], c3["b200"][2], c3["ref64"][2]
    assert set(g_ref) == set(g_b) == set(g_64)
    rows, zero = grad_report(g_b, g_ref, g_64)
    assert len(rows) > 100
    for direct, e_b, e_r, k in rows:
        assert direct < TOL or e_b <= 2 * max(e_r, TOL), (k, direct, e_b, e_r)
    for n_b, n_r, k in zero:  # structurally zero gradients stay noise in both arms
        assert n_b < 1e-5 and n_r < 1e-5, (k, n_b, n_r)
    print("max direct grad error %.2e; max err vs fp64: b200 %.2e, reference-fp32 %.2e" % tuple(
        max(r[i] for r in rows) for i in range(3)))


def test_s2_reference_engine_train_three_adam_steps(ref):
    """Engine.train (vision/train.py:120-157) itself -- zero_grad, prepare_mesh, Deformation, utils.chamfer_distance,
    backward, Adam -- over three batches: stock fp32, stock fp64 and installed.

    The first loss (no update yet) must agree to 1e-5.  From the second step on no two fp32 runs agree tightly: Adam's
    first updates are ~lr * sign(g), so wherever a gradient component is rounding noise the step is a coin flip, and
    the stock run is not even reproducible against itself (atomics in grid_sample's backward): measured here, the
    STOCK fp32 run is 0.1-0.6 % away from its own fp64 run at steps 2-3 and moves by as much between two launches.
    The installed path must stay within 4x the stock run's distance from fp64 or 2 %, whichever is larger, step by step;
    final weights within 4x the stock run's distance + a tenth of the Adam step budget."""
    import ref_harness as H
    args = H.c3_args(num_GCN_layers=6, hidden_GCN_size=120, number_points=2000)
    batches = [H.make_batch(args, B=2, seed=s) for s in (1, 2, 3)]
    info, mesh, net = ref.build(args, patched=False)
    state0 = {k: v.detach().clone() for k, v in net.state_dict().items()}
    steps_ref, mean_ref, final_ref = ref.engine_train(args, net, info, mesh, batches, patched=False)
    del net
    info64, mesh64, net64 = ref.build(args, patched=False, state=state0, dtype=torch.float64)
    steps_64, _, final_64 = ref.engine_train(args, net64, info64, mesh64, batches, patched=False)
    del net64
    info, mesh, net = ref.build(args, patched=True, state=state0)
    steps_b, mean_b, final_b = ref.engine_train(args, net, info, mesh, batches, patched=True)
    ptk_b200.uninstall()
    assert len(steps_ref) == len(steps_b) == len(steps_64) == 3
    assert abs(steps_b[0] - steps_ref[0]) < TOL * abs(steps_ref[0]), (steps_b, steps_ref)
    assert abs(mean_b - sum(steps_b) / 3) < 1e-4 * abs(mean_b)  # what Engine.train logged is what we observed
    for s_b, s_r, s_64 in zip(steps_b, steps_ref, steps_64):
        assert abs(s_b - s_64) <= max(4 * abs(s_r - s_64), 2e-2 * abs(s_64)), (steps_b, steps_ref, steps_64)
    lr_steps = args.lr * 3
    checked = 0
    for k, v64 in final_64.items():
        if not v64.is_floating_point() or "running" in k or "num_batches" in k:
            continue
        d64 = v64 - state0[k].double()
        if float(d64.abs().max()) < 0.1 * lr_steps:
            continue  # structurally zero gradient (conv bias before BatchNorm): pure noise in any fp32 run
        e_b = float((final_b[k].double() - v64).abs().max())
        e_r = float((final_ref[k].double() - v64).abs().max())
        assert e_b <= 4 * e_r + 0.1 * lr_steps, (k, e_b, e_r)
        checked += 1
    assert checked > 50
    print("Engine.train losses  b200 %s\n                     ref  %s\n                     fp64 %s" % (steps_b, steps_ref, steps_64))


# ----------------------------------------------------------------------------------------------- config 4
def test_best_step_batched_vs_the_references_candidate_loop(ref):
    """BASELINE config 4 as one call.  Reference arm = the loop of ActiveTouch.best_step (environment.py:167-180)
    written out: for each action, the STOCK Deformation under no_grad (compute_obs, :221-227), get_score =
    loss_coeff * chamfer (restated, :252-257), `.cpu()`, then `if s < best_score[e] and mask[e][i] == 0`.
    B200 arm = ptk_b200.policy.best_step_batched on the reference's Deformation after install() (tensor-core
    inference forward, fused sampling + Chamfer, arg-min on the device)."""
    import ref_harness as H
    args = H.c3_args(use_img=False, num_GCN_layers=8, hidden_GCN_size=200, number_points=2000)
    E, A, num = 3, 7, args.number_points
    g = torch.Generator().manual_seed(4)
    info, mesh, net = ref.build(args, patched=False)
    state = {k: v.detach().clone() for k, v in net.state_dict().items()}
    net.eval()
    batches = [[H.make_batch(args, B=1, seed=100 * e + a) for a in range(A)] for e in range(E)]
    touch = torch.stack([torch.stack([batches[e][a]["touch_charts"][0] for a in range(A)]) for e in range(E)]).cuda()
    touch = touch.view(E, A, -1, 4)                                    # (E, A, 125, 4)
    gt = torch.stack([batches[e][0]["gt_points"][0] for e in range(E)]).cuda()
    mask = torch.zeros(E, A)
    mask[0, 2] = 1
    mask[2, :] = 1                                                      # every action taken: the reference keeps None
    uniforms = [(torch.rand(E * A, num, generator=g).cuda(), torch.rand(2, E * A, num, generator=g).cuda())
                for _ in range(3)]
    vision = mesh[None].repeat(E, 1, 1)
    vmask = 3 * torch.ones(E, mesh.shape[0], 1, device="cuda")

    # ---- reference loop
    from oracle import torch_ref as tr
    best_actions, best_score = [None] * E, [1000] * E
    ref_scores = torch.zeros(E, A)
    for i in range(A):
        with torch.no_grad():
            charts = {"touch_charts": touch[:, i, :, :3], "touch_masks": touch[:, i, :, 3:],
                      "vision_charts": vision, "vision_masks": vmask}
            verts, _ = net(None, charts)
            cds = []
            rows = torch.arange(E) * A + i
            for uf, uv in uniforms:
                pts, _ = tr.batch_sample(verts, info["faces"], uf[rows], uv[:, rows])
                cds.append(tr.chamfer_distance(pts, gt)[0])
            score = (args.loss_coeff * torch.stack(cds).mean(0)).cpu()
        ref_scores[:, i] = score
        for e, s in enumerate(score):
            if s < best_score[e] and mask[e][i] == 0:
                best_actions[e], best_score[e] = i, s
    del net

    # ---- one call on the installed path
    info, mesh, net = ref.build(args, patched=True, state=state)
    net.eval()
    charts = {"touch_charts": touch[..., :3].contiguous(), "touch_masks": touch[..., 3:].contiguous(),
              "vision_charts": vision, "vision_masks": vmask}
    n0 = ptk_b200._lib.launch_count()
    actions, best, scores = ptk_b200.policy.best_step_batched(net, None, charts, gt, info["faces"], mask=mask.cuda(),
                                                              num=num, loss_coeff=args.loss_coeff, chunk=8,
                                                              uniforms=uniforms)
    assert ptk_b200._lib.launch_count() - n0 > 100, "the candidate pass must run libptk_b200 kernels"
    ptk_b200.uninstall()
    assert H.rel_err(scores, ref_scores) < 2e-5, H.rel_err(scores, ref_scores)
    want = [-1 if a is None else a for a in best_actions]
    assert actions.cpu().tolist() == want
    for e in range(E):
        if best_actions[e] is not None:
            assert abs(float(best[e]) - float(best_score[e])) <= 2e-5 * abs(float(best_score[e]))
        else:
            assert float(best[e]) == 1000.0
