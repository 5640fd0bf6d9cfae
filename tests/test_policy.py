"""Batched greedy-policy scoring (ptk_b200.policy) against the reference's per-action loop
(pterotactyl/policies/environment.py:167-213, 252-257)."""
import numpy as np
import pytest
import torch

import ptk_b200


def _reference_loop(scores, mask):
    """environment.py:170-180 restated on host lists: best_score starts at 1000, strict '<', masked actions skipped."""
    E, A = scores.shape
    best_a, best_s = [None] * E, [1000.0] * E
    for i in range(A):
        for e in range(E):
            if scores[e, i] < best_s[e] and mask[e][i] == 0:
                best_a[e], best_s[e] = i, float(scores[e, i])
    return best_a, best_s


def test_best_actions_matches_reference_loop_cpu():
    rng = np.random.default_rng(0)
    scores = torch.from_numpy(rng.random((6, 50)).astype(np.float32) * 100)
    scores[1, 7] = scores[1, 3] = scores[1].min() - 1          # exact tie: the lower action must win
    scores[2] = 2000.0                                          # nothing below 1000: no action
    mask = torch.zeros(6, 50, dtype=torch.int64)
    mask[0, int(scores[0].argmin())] = 1                        # the best action was already taken
    mask[3] = 1                                                 # everything taken
    a, s = ptk_b200.policy.best_actions(scores, mask)
    ra, rs = _reference_loop(scores.numpy(), mask.numpy())
    assert a.tolist() == [-1 if x is None else x for x in ra]
    assert np.allclose(s.numpy(), np.array(rs, np.float32))


@pytest.mark.gpu
def test_score_candidates_equals_per_action_calls(golden):
    m = golden("meshes")
    verts0 = torch.from_numpy(m["obj0_verts"]).cuda()
    faces = torch.from_numpy(m["obj0_faces"].astype(np.int64)).cuda()
    E, A, num = 3, 5, 1500
    g = torch.Generator(device="cuda").manual_seed(0)
    verts = verts0[None, None] * (1.0 + 0.05 * torch.rand(E, A, 1, 1, device="cuda", generator=g))
    gt = ptk_b200.utils.batch_sample(verts0[None].repeat(E, 1, 1), faces, num=2000, generator=g)
    uniforms = [ptk_b200.utils.draw_uniforms(E * A, num, "cuda", g) for _ in range(3)]
    scores = ptk_b200.policy.score_candidates(verts, faces, gt, num=num, uniforms=uniforms)
    assert scores.shape == (E, A) and not scores.requires_grad
    # the reference's loop: one utils.chamfer_distance call per action over the environment batch
    for a in range(A):
        rows = torch.arange(E, device="cuda") * A + a
        uni = [(uf[rows].contiguous(), uv[:, rows].contiguous()) for uf, uv in uniforms]
        want = 9000.0 * ptk_b200.utils.chamfer_distance(verts[:, a].contiguous(), faces, gt, num=num, uniforms=uni)
        assert torch.equal(scores[:, a], want)          # batching changes nothing, bit for bit
        # ... and the score itself against the reference's get_score restated with eager torch ops
        # (environment.py:252-257 -> utils.py:204-217 over oracle/torch_ref.py), same uniforms
        from oracle import torch_ref as tr
        cds = []
        for uf, uv in uni:
            pts, _ = tr.batch_sample(verts[:, a].contiguous(), faces, uf, uv)
            cds.append(tr.chamfer_distance(pts, gt)[0])
        ref_score = 9000.0 * torch.stack(cds).mean(0)
        assert float((scores[:, a] - ref_score).abs().max() / ref_score.abs().max()) < 1e-5
    act, best = ptk_b200.policy.best_actions(scores, torch.zeros(E, A, device="cuda"))
    assert torch.equal(act, scores.argmin(1)) and torch.equal(best, scores.min(1).values)


@pytest.mark.gpu
@pytest.mark.parametrize("B,N,C", [(1, 1, 1), (3, 1949, 300), (2, 2324, 50), (5, 37, 100)])
def test_vertex_max_matches_torch(B, N, C):
    g = torch.Generator(device="cuda").manual_seed(B + N + C)
    x = torch.randn(B, N, C, device="cuda", generator=g)
    x[:, N // 2] = x[:, 0]  # ties: the lower vertex id wins
    xr = x.clone().requires_grad_(True)
    xo = x.clone().requires_grad_(True)
    v, arg = ptk_b200.ops.vertex_max(xo)
    want = xr.max(dim=1)[0]
    assert torch.equal(v, want)
    first = (xr == want[:, None, :]).to(torch.int32).argmax(dim=1)
    assert torch.equal(arg.long(), first.long())
    w = torch.rand(B, C, device="cuda", generator=g)
    (v * w).sum().backward()
    ref = torch.zeros_like(x)
    ref.scatter_(1, first[:, None, :].long(), w[:, None, :])
    assert torch.equal(xo.grad, ref)


@pytest.mark.gpu
def test_vertex_max_nan_propagates():
    x = torch.randn(2, 50, 8, device="cuda")
    x[1, 20, 3] = float("nan")
    v, _ = ptk_b200.ops.vertex_max(x)
    assert torch.isnan(v[1, 3]) and not torch.isnan(v[0]).any() and int(torch.isnan(v).sum()) == 1
