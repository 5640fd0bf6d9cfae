"""Batched candidate scoring for the greedy / oracle touch policies.

The reference scores candidate actions one at a time: `ActiveTouch.best_step`
(pterotactyl/policies/environment.py:167-213) loops over `num_actions` (50), and for every action runs
the deformation, `get_score` = `loss_coeff * utils.chamfer_distance(verts, faces, gt, num=number_points)`
followed by `.cpu()` (environment.py:252-257), then compares scores per environment on the host
(environment.py:176-180).  The candidates are independent, so here all `E * A` candidate meshes are
scored by ONE sample + Chamfer pass and the masked arg-min runs on the device; nothing is read back
per candidate.  The simulator, the touch CNN and the deformation encoders stay the reference's own
code: this module starts from the candidate meshes they produce.

`best_step_batched` is the whole inner loop of `best_step` in one call: deformation network (no grad) on all
E * A candidates -> 3 x (surface sampling + Chamfer) -> masked arg-min, in device-sized chunks.

Multi-GPU: candidates are whole objects, so `score_candidates(..., shard=True)` / `best_step_batched(...,
shard=True)` split the E*A rows over the ranks (ptk_b200.dist) and all-gather the (E*A,) score vector -- the
only collective.
"""
import torch

from . import dist as _dist
from . import utils


def score_candidates(verts, faces, gt_points, num=10000, loss_coeff=9000.0, repeat=3, generator=None,
                     uniforms=None, shard=False):
    """verts (E, A, V, 3) candidate meshes (one per environment and action), faces (F,3) shared,
    gt_points (E, P, 3) -> scores (E, A) = loss_coeff * chamfer_distance, on the device, no grad."""
    if verts.dim() != 4 or verts.shape[-1] != 3:
        raise ValueError(f"verts must be (E, A, V, 3), got {tuple(verts.shape)}")
    E, A, V, _ = verts.shape
    if gt_points.shape[0] != E:
        raise ValueError("gt_points must have one cloud per environment")
    with torch.no_grad():
        flat = verts.reshape(E * A, V, 3)
        lo, hi = 0, E * A
        if shard:
            rank, world = _dist.rank_world()
            lo, hi = _dist.shard_bounds(E * A, rank, world)
        env = torch.arange(lo, hi, device=verts.device) // A
        gt = gt_points.index_select(0, env)  # every candidate of an environment is compared with its cloud
        uni = None
        if uniforms is not None:  # [(u_face (E*A,num), uv (2,E*A,num))] * repeat for the full candidate list
            uni = [(uf[lo:hi].contiguous(), uv[:, lo:hi].contiguous()) for uf, uv in uniforms]
        cd = utils.chamfer_distance(flat[lo:hi].contiguous(), faces, gt, num=num, repeat=repeat,
                                    generator=generator, uniforms=uni)
        scores = loss_coeff * cd
        if shard:
            scores = _dist.gather_objects_vector(scores, E * A)
        return scores.reshape(E, A)


def best_actions(scores, mask=None):
    """Per-environment greedy choice of environment.py:176-180 on the device: the lowest score among the
    actions whose mask entry is 0; strict '<' in ascending action order => the lowest action index wins
    ties; -1 (the reference keeps None) when every action is masked or no score is below 1000.
    scores (E, A), mask (E, A) or None -> (action (E,) int64, score (E,))."""
    s = scores.clone()
    if mask is not None:
        s = torch.where(mask.to(s.device) != 0, torch.full_like(s, float("inf")), s)
    s = torch.where(s < 1000.0, s, torch.full_like(s, float("inf")))  # best_score starts at 1000 (environment.py:170)
    best, arg = s.min(dim=1)
    # torch.min may return any of several equal minima: take the first one explicitly
    first = (s == best[:, None]).to(torch.int64).argmax(dim=1)
    arg = torch.where(torch.isinf(best), torch.full_like(first, -1), first)
    return arg, torch.where(torch.isinf(best), torch.full_like(best, 1000.0), best)


def best_step_batched(deform, img, charts, gt_points, faces, mask=None, num=10000, loss_coeff=9000.0, repeat=3,
                      chunk=400, shard=False, generator=None, uniforms=None):
    """The candidate loop of ActiveTouch.best_step (environment.py:167-180) as one batched pass.

    The reference evaluates, for each of the A actions in turn, `compute_obs` = `self.deform(img, charts)` under
    no_grad followed by `get_score` = `loss_coeff * utils.chamfer_distance(verts, faces, gt, num)` and `.cpu()`
    (environment.py:221-257), then keeps per environment the lowest score among unmasked actions.  Here:

    deform     the deformation network, called as the reference calls it: `deform(img, charts) -> (verts, mask)`
               (the reference's own `Deformation` after ptk_b200.install(), or any callable with that contract)
    img        (E, 3, H, W) one image per environment (or None for touch-only models); repeated per candidate
    charts     the dict `prepare_mesh` / `get_inputs` build, with a leading (E, A) candidate grid on the touch
               entries: 'touch_charts' (E, A, T, 3), 'touch_masks' (E, A, T, 1); 'vision_charts' (E, V0, 3) and
               'vision_masks' (E, V0, 1) are per environment (optionally (E, A, ...) too)
    gt_points  (E, P, 3); faces (F, 3) shared; mask (E, A) nonzero = action already taken
    chunk      candidates per deformation pass (bounds activation memory: 400 x 1949 x 448 floats = 1.4 GB)

    Returns (actions (E,) int64 with -1 where the reference keeps None, best scores (E,), scores (E, A)).
    With shard=True every rank evaluates its contiguous block of the E*A candidates and the scores are
    all-gathered, so all ranks return the same decision."""
    tc = charts["touch_charts"]
    if tc.dim() != 4:
        raise ValueError(f"touch_charts must be (E, A, T, 3), got {tuple(tc.shape)}")
    E, A = tc.shape[:2]
    if gt_points.shape[0] != E:
        raise ValueError("gt_points must have one cloud per environment")
    lo, hi = 0, E * A
    if shard:
        rank, world = _dist.rank_world()
        lo, hi = _dist.shard_bounds(E * A, rank, world)

    def cand(t, idx):
        """Rows `idx` of a chart tensor: per candidate when it carries the (E, A) grid, else per environment."""
        if t.dim() == 4 and tuple(t.shape[:2]) == (E, A):
            return t.reshape(E * A, *t.shape[2:]).index_select(0, idx)
        return t.index_select(0, idx // A)

    env = lambda t, idx: t.index_select(0, idx // A)  # img, gt_points: one per environment

    out = []
    with torch.no_grad():
        for s in range(lo, hi, max(int(chunk), 1)):
            e = min(s + max(int(chunk), 1), hi)
            idx = torch.arange(s, e, device=tc.device)
            sub = {k: cand(v, idx) for k, v in charts.items()}
            verts = deform(env(img, idx) if img is not None else None, sub)
            verts = verts[0] if isinstance(verts, (tuple, list)) else verts
            uni = None
            if uniforms is not None:
                uni = [(uf[s:e].contiguous(), uv[:, s:e].contiguous()) for uf, uv in uniforms]
            cd = utils.chamfer_distance(verts, faces, env(gt_points, idx), num=num, repeat=repeat,
                                        generator=generator, uniforms=uni)
            out.append(loss_coeff * cd)
        scores = torch.cat(out) if out else torch.empty(0, device=tc.device)
        if shard:
            scores = _dist.gather_objects_vector(scores, E * A)
        scores = scores.reshape(E, A)
        actions, best = best_actions(scores, mask)
    return actions, best, scores
