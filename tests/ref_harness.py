"""TEST INFRASTRUCTURE: drives the reference's own, unedited modules (an installed / checked-out `pterotactyl`,
located by ptk_b200.find_reference) once as the reference runs them -- dense adjacency, torch GEMMs, its own
GCN / Deformation classes, Chamfer loss from oracle/torch_ref.py -- and once with ptk_b200.install() applied, on
the same GPU, the same seeds, the same inputs.  Used by tests/test_reference_gpu.py and tools/reference_step.py.

The loss of the reference arm is the reference's OWN `utils.chamfer_distance` body (Tensor.multinomial, torch.rand, its
gathers) with only the PyTorch3D names rebound to the eager-torch restatement (PyTorch3D is absent).
install() keeps the reference's RNG stream (face_draw="multinomial"), so both arms draw the same samples from the same
seed.
"""
import copy
import os
import types

import numpy as np
import torch

import ptk_b200
from oracle import torch_ref as tr


def rel_err(a, b):
    a = np.asarray(a.detach().double().cpu() if torch.is_tensor(a) else a, np.float64)
    b = np.asarray(b.detach().double().cpu() if torch.is_tensor(b) else b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def grad_report(g_b, g_ref, g_64):
    """Per parameter: (direct error b200 vs reference-fp32, b200 vs fp64, reference-fp32 vs fp64, name).
    Parameters whose fp64 gradient is structurally zero (conv biases in front of a BatchNorm, GCN bias entries
    beyond the propagated slice: < 1e-9 of the largest gradient) are reported apart as absolute noise / scale."""
    scale = max(float(v.abs().max()) for v in g_64.values())
    rows, zero = [], []
    for k in g_ref:
        if float(g_64[k].abs().max()) < 1e-9 * scale:
            zero.append((float(g_b[k].abs().max()) / scale, float(g_ref[k].abs().max()) / scale, k))
        else:
            rows.append((rel_err(g_b[k], g_ref[k]), rel_err(g_b[k], g_64[k]), rel_err(g_ref[k], g_64[k]), k))
    return sorted(rows, reverse=True), sorted(zero, reverse=True)


def strict_fp32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False


def c3_args(**over):
    """BASELINE config 3 flags (vision/train.py:287-401 defaults) -- `v_t_p`."""
    a = dict(use_img=True, use_touch=True, finger=True, num_grasps=5, num_GCN_layers=20, hidden_GCN_size=300, cut=0.33,
             num_CNN_blocks=6, layers_per_block=3, CNN_ker_size=5, number_points=3000, loss_coeff=9000.0, lr=3e-4,
             batch_size=2, seed=0, eval=False, exp_id="ptk", exp_type="ptk", epochs=1)
    a.update(over)
    return types.SimpleNamespace(**a)


def make_batch(args, B, seed=0, n_gt=None):
    """One synthetic batch in the layout of data_loaders.mesh_loader_vision.collate: image,
    touch charts (B, fingers*grasps, 25, 4) = xyz + mask in {0,1,2} (environment.py:304-315), ground-truth cloud."""
    g = torch.Generator().manual_seed(seed)
    k = (1 if args.finger else 4) * args.num_grasps
    n_gt = n_gt or args.number_points
    img = torch.rand(B, 3, 256, 256, generator=g)
    centre = torch.nn.functional.normalize(torch.randn(B, k, 1, 3, generator=g), dim=-1) * 0.25
    offs = (torch.rand(B, k, 25, 3, generator=g) - 0.5) * 0.017
    mask = torch.multinomial(torch.tensor([0.2, 0.2, 0.6]), B * k, replacement=True, generator=g).view(B, k, 1, 1)
    xyz = torch.where(mask == 2, centre + offs, torch.where(mask == 1, centre.expand(-1, -1, 25, -1),
                                                            torch.zeros(B, k, 25, 3)))
    touch = torch.cat((xyz, mask.float().expand(-1, -1, 25, 1)), dim=-1)
    d = torch.nn.functional.normalize(torch.randn(B, n_gt, 3, generator=g), dim=-1)
    gt = d * (0.25 + 0.03 * torch.sin(7 * d[..., :1]))
    return {"img": img, "touch_charts": touch, "gt_points": gt.contiguous()}


def packed_areas(V, F):
    """pytorch3d.ops.mesh_face_areas_normals restated in eager torch (packed verts / faces): (areas, None)."""
    v0, v1, v2 = V[F[:, 0]], V[F[:, 1]], V[F[:, 2]]
    a, b = v1 - v0, v2 - v0
    cx = a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1]
    cy = a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2]
    cz = a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]
    return torch.sqrt((cx * cx + cy * cy) + cz * cz) * 0.5, None


class restated_pytorch3d:
    """Context manager: the reference's `utils` module with its PyTorch3D names rebound to the eager-torch
    restatement (oracle/torch_ref.py) -- i.e. the reference's own `batch_sample` / `chamfer_distance` BODIES
    (Tensor.multinomial, torch.rand, gathers: utils.py:152-217) over brute-force nearest neighbours.  This is the
    reference arm's loss; PyTorch3D itself is not installable."""

    def __init__(self, ref_utils):
        self.U = ref_utils

    def __enter__(self):
        U = self.U
        assert U.chamfer_distance.__module__ == "pterotactyl.utility.utils", "uninstall() first"
        self.saved = (U.cuda_cd, U.mesh_face_areas_normals)
        U.cuda_cd = lambda x, y, batch_reduction=None: (tr.chamfer_autograd(x, y), None)
        U.mesh_face_areas_normals = packed_areas
        return U

    def __exit__(self, *exc):
        self.U.cuda_cd, self.U.mesh_face_areas_normals = self.saved


class Reference:
    """The reference's modules, imported unedited."""

    def __init__(self):
        self.utils, self.vision_model = ptk_b200.import_reference("utility.utils", "reconstruction.vision.model")
        self.root = ptk_b200.find_reference()
        self.objects = os.path.join(self.root, "pterotactyl", "objects")
        ptk_b200.utils.set_object_dir(self.objects)

    def chart(self, name="vision_charts.obj"):
        return os.path.join(self.objects, name)

    # ------------------------------------------------------------------ the two arms
    def build(self, args, patched, state=None, dtype=torch.float32):
        """(mesh_info, initial_mesh, Deformation) through the reference's own constructors (vision/train.py:51-62).
        patched=True: after ptk_b200.install(); patched=False: the stock classes, dense adjacency."""
        ptk_b200.uninstall()
        if patched:
            import importlib
            mods = [self.vision_model]
            for n in ("reconstruction.autoencoder.model", "policies.DDQN.model"):
                mods.append(importlib.import_module("pterotactyl." + n))
            ptk_b200.install(self.utils, mods)
        mesh_info, initial_mesh = self.utils.load_mesh_vision(args, self.chart())
        initial_mesh = initial_mesh.cuda()
        torch.manual_seed(args.seed)
        net = self.vision_model.Deformation(mesh_info, initial_mesh, args)
        net.cuda()
        if state is not None:
            net.load_state_dict(state)
        if dtype != torch.float32:
            net = net.to(dtype)
            mesh_info = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in mesh_info.items()}
            net.adj_info = mesh_info
            for enc in (net.img_encoder_global, net.img_encoder_local) if args.use_img else ():
                enc.matrix = enc.matrix.to(dtype)
            initial_mesh = initial_mesh.to(dtype)
        return mesh_info, initial_mesh, net

    def step(self, args, net, mesh_info, initial_mesh, batch, patched, seed=123):
        """Loop body of Engine.train (vision/train.py:122-147) without the optimizer: returns verts, loss, grads."""
        vm = self.vision_model
        dt = next(net.parameters()).dtype
        net.train()
        net.zero_grad(set_to_none=True)
        img = batch["img"].cuda().to(dt)
        gt_points = batch["gt_points"].cuda().to(dt)
        with torch.no_grad():
            charts = vm.prepare_mesh(batch, initial_mesh, args)
            charts = {k: v.to(dt) for k, v in charts.items()}
        verts = net(img, charts)[0]
        torch.manual_seed(seed)
        if patched:  # install() put ptk_b200's function there; it draws faces like the reference (multinomial)
            assert self.utils.chamfer_distance is ptk_b200.utils.chamfer_distance
            loss = self.utils.chamfer_distance(verts, mesh_info["faces"], gt_points, num=args.number_points)
        else:        # the reference's own function body
            with restated_pytorch3d(self.utils) as U:
                loss = U.chamfer_distance(verts, mesh_info["faces"], gt_points, num=args.number_points)
        loss = args.loss_coeff * loss.mean()
        loss.backward()
        grads = {k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None}
        return verts.detach(), loss.detach(), grads

    def engine_train(self, args, net, mesh_info, initial_mesh, batches, patched, seed=321):
        """The reference's Engine.train METHOD, unedited (vision/train.py:120-157), on a list of batches.
        Returns (per-step losses, mean loss it logged, final state dict).  A float64 `net` gets float64 batches."""
        train_mod = ptk_b200.import_reference("reconstruction.vision.train")
        dt = next(net.parameters()).dtype
        if dt != torch.float32:
            batches = [{k: (v.to(dt) if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in b.items()}
                       for b in batches]
        eng = train_mod.Engine.__new__(train_mod.Engine)  # __init__ makes directories and needs a config dump
        eng.args, eng.encoder, eng.mesh_info, eng.initial_mesh = args, net, mesh_info, initial_mesh
        eng.epoch, eng.best_loss = 0, 10000
        eng.optimizer = torch.optim.Adam(list(net.parameters()), lr=args.lr, weight_decay=0)
        logged, steps = {}, []
        writer = types.SimpleNamespace(add_scalars=lambda tag, vals, epoch: logged.update({tag: dict(vals)}))
        saved = inner = train_mod.utils.chamfer_distance
        if patched:
            assert inner is ptk_b200.utils.chamfer_distance, "install() did not reach the trainer's utils module"
        else:
            assert inner.__module__ == "pterotactyl.utility.utils"

        def observed(verts, faces, gt_points, num=1000, repeat=3):  # same call, the per-object distances recorded
            cd = inner(verts, faces, gt_points, num=num, repeat=repeat)
            steps.append(float(args.loss_coeff * cd.detach().double().mean()))
            return cd

        import contextlib
        ctx = contextlib.nullcontext() if patched else restated_pytorch3d(train_mod.utils)
        with ctx:
            train_mod.utils.chamfer_distance = observed
            try:
                torch.manual_seed(seed)
                eng.train(batches, writer)
            finally:
                train_mod.utils.chamfer_distance = saved
        return steps, logged["train_loss"][args.exp_id], copy.deepcopy(net.state_dict())
