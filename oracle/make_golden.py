"""Generates tests/golden/*.npz from the REFERENCE ITSELF (run in the dev container, where
/root/reference exists; the GPU box only sees the committed fixtures).

    python oracle/make_golden.py

What is pinned, and against what:
  * meshes.npz        -- vertices/faces of the reference's shipped .obj assets (inputs only).
  * adjacency.npz     -- CSR of adj_info['origional'] / ['adj'] produced by the reference's OWN
                         utils.adj_init (pterotactyl/utility/utils.py:56-71) for the vision-only,
                         finger ('p', N=1949) and grasp ('g', N=2324) settings.
  * gcn.npz           -- outputs and gradients of the reference's OWN GCN / GCN_layer
                         (pterotactyl/reconstruction/vision/model.py:290-363) on those adjacencies.
  * chamfer.npz       -- BASELINE config 1 (two 10k clouds from objects/test_objects/0.obj) and
                         tie / P1!=P2 cases through a torch restatement of PyTorch3D's
                         knn_points/chamfer_distance (PyTorch3D is not vendored: "parity unpinned"
                         against PyTorch3D itself, pinned against this independent restatement).
  * sampler.npz       -- the reference's OWN utils.batch_sample / utils.chamfer_distance run with
                         pytorch3d.* stubbed by oracle/torch_ref.py and torch.multinomial replaced by
                         the explicit-uniform face pick, incl. degenerate meshes.

The reference modules are imported unmodified; missing third-party imports (matplotlib, pytorch3d,
trimesh, pyrender) are stubbed in sys.modules and `.cuda()` is made a no-op (no GPU here).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("PTK_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import oracle as orc  # noqa: E402
from oracle import torch_ref as tr  # noqa: E402


def parse_obj(path):
    vs, fs = [], []
    with open(path) as fh:
        for line in fh:
            t = line.split()
            if not t:
                continue
            if t[0] == "v":
                vs.append([float(t[1]), float(t[2]), float(t[3])])
            elif t[0] == "f":
                idx = [int(tok.split("/")[0]) for tok in t[1:]]
                idx = [i - 1 if i > 0 else len(vs) + i for i in idx]
                for k in range(1, len(idx) - 1):
                    fs.append([idx[0], idx[k], idx[k + 1]])
    return np.asarray(vs, np.float32), np.asarray(fs, np.int64)


def install_stubs(meshes):
    """Stub the absent third-party modules, then import the reference's utils + model."""
    for name in ["matplotlib", "matplotlib.pyplot", "pytorch3d", "pytorch3d.loss", "pytorch3d.ops",
                 "pytorch3d.ops.mesh_face_areas_normals", "pytorch3d.ops.sample_points_from_meshes",
                 "pytorch3d.io", "pytorch3d.io.obj_io", "trimesh", "pyrender", "pybullet"]:
        sys.modules.setdefault(name, types.ModuleType(name))

    state = {}

    def cuda_cd(x, y, batch_reduction=None):
        cham, _, _ = tr.chamfer_distance(x, y)
        return cham, None

    def mesh_face_areas_normals(V, F):
        v0, v1, v2 = V[F[:, 0]], V[F[:, 1]], V[F[:, 2]]
        a, b = v1 - v0, v2 - v0
        cx = a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1]
        cy = a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2]
        cz = a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]
        areas = tr.ieee_sqrt((cx * cx + cy * cy) + cz * cz) * 0.5
        state["areas"] = areas.clone()
        return areas, None

    def _rand_barycentric_coords(s1, s2, dtype, device):
        return tr.rand_barycentric_from(state["uv"])

    def load_obj(path):
        v, f = parse_obj(path)
        faces = types.SimpleNamespace(verts_idx=torch.from_numpy(f))
        return torch.from_numpy(v), faces, None

    sys.modules["pytorch3d.loss"].chamfer_distance = cuda_cd
    sys.modules["pytorch3d.ops.mesh_face_areas_normals"].mesh_face_areas_normals = mesh_face_areas_normals
    sys.modules["pytorch3d.ops.sample_points_from_meshes"]._rand_barycentric_coords = _rand_barycentric_coords
    sys.modules["pytorch3d.io.obj_io"].load_obj = load_obj
    sys.modules["pytorch3d.io.obj_io"].save_obj = None
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self

    from pterotactyl.utility import utils as ref_utils
    from pterotactyl.reconstruction.vision import model as ref_model

    # torch.multinomial consumes an implementation-defined RNG stream; the explicit-uniform pick
    # replaces exactly that one call (utils.py:170) -- everything around it is the reference's code.
    # The integer weights are defined on the RAW areas (a_f / a_max), which the areas stub stashes:
    # the reference's a_f / sum(a) normalisation (utils.py:166) rescales every weight by the same
    # factor and so leaves the distribution unchanged, but would add one more fp32 rounding.
    def multinomial(self, num, replacement=True):
        cum = tr.face_cumweights(state["areas"].reshape(self.shape).numpy())
        return torch.from_numpy(tr.pick_faces(cum, state["u_face"].numpy()))

    torch.Tensor.multinomial = multinomial
    return ref_utils, ref_model, state


def main():
    os.makedirs(OUT, exist_ok=True)
    obj_dir = os.path.join(REF, "pterotactyl", "objects")
    meshes = {}
    for key, rel in [("vision", "vision_charts.obj"), ("touch", "touch_chart.obj"),
                     ("obj0", "test_objects/0.obj"), ("obj1", "test_objects/1.obj")]:
        v, f = parse_obj(os.path.join(obj_dir, rel))
        meshes[key + "_verts"], meshes[key + "_faces"] = v, f.astype(np.int32)
        print(key, v.shape, f.shape)
    np.savez_compressed(os.path.join(OUT, "meshes.npz"), **meshes)

    ref_utils, ref_model, state = install_stubs(meshes)

    # ------------------------------------------------------------ adjacency from the reference
    adj_out = {}
    dense = {}
    vision_obj = os.path.join(obj_dir, "vision_charts.obj")
    for tag, use_touch, finger in [("v", False, False), ("p", True, True), ("g", True, False)]:
        args = types.SimpleNamespace(use_touch=use_touch, finger=finger, num_grasps=5)
        adj_info, verts = ref_utils.load_mesh_vision(args, vision_obj)
        for which in ("origional", "adj"):
            rp, col = tr.dense_to_csr(adj_info[which])
            adj_out[f"{tag}_{which}_rowptr"], adj_out[f"{tag}_{which}_col"] = rp, col
            dense[(tag, which)] = adj_info[which]
        adj_out[f"{tag}_faces"] = adj_info["faces"].numpy().astype(np.int32)
        print("adj", tag, adj_info["adj"].shape, "nnz", len(adj_out[f"{tag}_adj_col"]),
              "faces", adj_info["faces"].shape)
    np.savez_compressed(os.path.join(OUT, "adjacency.npz"), **adj_out)

    # ------------------------------------------------------------ reference GCN
    gcn_out = {}
    cases = [
        # name, adjacency, in, hidden, layers, cut, B, ignore_touch_matrix
        ("p_small", ("p", "adj"), 16, 40, 3, 0.33, 2, False),
        ("g_small", ("g", "adj"), 12, 36, 3, 0.33, 1, False),
        ("v_orig", ("p", "origional"), 16, 100, 4, 0.33, 1, True),
        ("p_default", ("p", "adj"), 50, 300, 20, 0.33, 1, False),
    ]
    for name, akey, cin, hid, nl, cut, B, ignore in cases:
        torch.manual_seed(1234)
        args = types.SimpleNamespace(num_GCN_layers=nl, hidden_GCN_size=hid, cut=cut)
        net = ref_model.GCN(cin, args, ignore_touch_matrix=ignore)
        adj = dense[akey]
        N = adj.shape[0]
        x = torch.rand(B, N, cin)
        x.requires_grad_(True)
        info = {"origional": dense[(akey[0], "origional")], "adj": dense[(akey[0], "adj")]}
        y = net(x, info)
        gout = torch.rand(B, N, 3, generator=torch.Generator().manual_seed(7))
        (y * gout).sum().backward()
        gcn_out[name + "_meta"] = np.array([cin, hid, nl, B, N, int(ignore)], np.int64)
        gcn_out[name + "_cut"] = np.array([cut], np.float64)
        gcn_out[name + "_y"] = y.detach().numpy()
        gcn_out[name + "_gx"] = x.grad.numpy() if name != "p_default" else x.grad.numpy()[:, ::16]
        if name != "p_default":
            gcn_out[name + "_x"] = x.detach().numpy()
            gcn_out[name + "_gout"] = gout.numpy()
            for i, layer in enumerate(net.layers):
                gcn_out[f"{name}_w{i}"] = layer.weight.detach().numpy()
                gcn_out[f"{name}_b{i}"] = layer.bias.detach().numpy()
                gcn_out[f"{name}_gw{i}"] = layer.weight.grad.numpy()
                gcn_out[f"{name}_gb{i}"] = layer.bias.grad.numpy()
        else:
            # full default 20x300 net: weights/inputs are regenerated from the seed by the test
            # (same torch build in the image), only outputs and a few gradients are stored.
            gcn_out[name + "_gw0"] = net.layers[0].weight.grad.numpy()
            gcn_out[name + "_gb0"] = net.layers[0].bias.grad.numpy()
            gcn_out[name + "_gw19"] = net.layers[19].weight.grad.numpy()
            gcn_out[name + "_w5_sum"] = np.array([net.layers[5].weight.double().sum().item()])
        print("gcn", name, y.shape, float(y.abs().mean()))
    np.savez_compressed(os.path.join(OUT, "gcn.npz"), **gcn_out)

    # ------------------------------------------------------------ sampler via the reference's batch_sample
    samp = {}
    g = torch.Generator().manual_seed(0)
    v0 = torch.from_numpy(meshes["obj0_verts"])
    f0 = torch.from_numpy(meshes["obj0_faces"].astype(np.int64))
    vis_v = torch.from_numpy(meshes["vision_verts"])
    touch_v = torch.from_numpy(meshes["touch_verts"])
    p_faces = torch.from_numpy(adj_out["p_faces"].astype(np.int64))

    def run_sample(tag, verts, faces, S, seed):
        gg = torch.Generator().manual_seed(seed)
        B = verts.shape[0]
        state["u_face"] = torch.rand(B, S, generator=gg)
        state["uv"] = torch.rand(2, B, S, generator=gg)
        vv = verts.clone().requires_grad_(True)
        pts = ref_utils.batch_sample(vv, faces.clone(), num=S)
        gp = torch.rand(B, S, 3, generator=gg)
        (pts * gp).sum().backward()
        samp[tag + "_verts"] = verts.numpy()
        samp[tag + "_faces"] = faces.numpy().astype(np.int32)
        samp[tag + "_u_face"] = state["u_face"].numpy()
        samp[tag + "_uv"] = state["uv"].numpy()
        samp[tag + "_pts"] = pts.detach().numpy()
        samp[tag + "_gpts"] = gp.numpy()
        samp[tag + "_gverts"] = vv.grad.numpy()
        # cross-check against the C oracle right here
        opts, ofi = orc.sample_fwd(verts.numpy(), faces.numpy(), state["u_face"].numpy(), state["uv"].numpy())
        assert np.array_equal(opts, pts.detach().numpy()), tag + ": C oracle != reference batch_sample"
        print("sample", tag, pts.shape, "C oracle bit-exact")

    # (a) test object 0, B=2 (second copy scaled)
    run_sample("obj0", torch.stack([v0, v0 * 1.5 + 0.01]), f0, 2000, 11)
    # (b) fused vision + 5 touch charts ('p' faces); touch charts: real / collapsed (mask 1) / zeros (mask 0)
    tc = []
    for i in range(5):
        if i % 3 == 0:
            tc.append(touch_v + torch.tensor([0.2, 0.05 * i, 0.1]))
        elif i % 3 == 1:
            tc.append(torch.ones(25, 3) * 0.123)  # repeated point -> zero-area faces
        else:
            tc.append(torch.zeros(25, 3))
    vp = torch.cat([vis_v] + tc)
    run_sample("p_mesh", torch.stack([vp, vp * 0.5]), p_faces, 3000, 12)
    # (c) all-degenerate mesh -> uniform fallback (utils.py:166-168)
    run_sample("degenerate", torch.ones(1, 25, 3) * 0.5,
               torch.from_numpy(meshes["touch_faces"].astype(np.int64)), 500, 13)
    # (d) the touch chart (F=32), config-2 shape family
    run_sample("touch", torch.stack([touch_v, touch_v * 2.0, touch_v + 0.3]),
               torch.from_numpy(meshes["touch_faces"].astype(np.int64)), 4000, 14)

    # utils.chamfer_distance (repeat=3, mean over repeats) on the reference code path
    B, S = 2, 1500
    gg = torch.Generator().manual_seed(21)
    verts = torch.stack([v0, v0 * 1.2])
    gt = torch.rand(B, 1200, 3, generator=gg) * 0.2 - 0.1
    us = [(torch.rand(B, S, generator=gg), torch.rand(2, B, S, generator=gg)) for _ in range(3)]
    calls = {"n": 0}
    orig_bs = ref_utils.batch_sample

    def bs_hook(v, f, num=10000):
        state["u_face"], state["uv"] = us[calls["n"]]
        calls["n"] += 1
        return orig_bs(v, f, num=num)

    ref_utils.batch_sample = bs_hook
    vv = verts.clone().requires_grad_(True)
    # the reference's chamfer (knn) is non-differentiable in the stub; use the autograd restatement
    sys.modules["pytorch3d.loss"].chamfer_distance = None
    ref_utils.cuda_cd = lambda x, y, batch_reduction=None: (tr.chamfer_autograd(x, y), None)
    cd = ref_utils.chamfer_distance(vv, f0, gt, num=S, repeat=3)
    cd.sum().backward()
    ref_utils.batch_sample = orig_bs
    samp["meshcd_verts"] = verts.numpy()
    samp["meshcd_faces"] = f0.numpy().astype(np.int32)
    samp["meshcd_gt"] = gt.numpy()
    for r in range(3):
        samp[f"meshcd_u_face{r}"] = us[r][0].numpy()
        samp[f"meshcd_uv{r}"] = us[r][1].numpy()
    samp["meshcd_cd"] = cd.detach().numpy()
    samp["meshcd_gverts"] = vv.grad.numpy()
    print("mesh chamfer", cd.detach().numpy())
    np.savez_compressed(os.path.join(OUT, "sampler.npz"), **samp)

    # ------------------------------------------------------------ Chamfer goldens
    ch = {}
    # config 1: two 10k clouds from test_objects/0.obj, seeds 0 and 1
    clouds = []
    for seed in (0, 1):
        gg = torch.Generator().manual_seed(seed)
        uf = torch.rand(1, 10000, generator=gg)
        uv = torch.rand(2, 1, 10000, generator=gg)
        pts, _ = orc.sample_fwd(v0[None].numpy(), f0.numpy(), uf.numpy(), uv.numpy())
        clouds.append(torch.from_numpy(pts))
    x, y = clouds
    cham, ix, iy = tr.chamfer_distance(x, y)
    ch["c1_x"], ch["c1_y"] = x.numpy(), y.numpy()
    ch["c1_cham"] = cham.numpy()
    ch["c1_idx_x"], ch["c1_idx_y"] = ix.numpy().astype(np.int32), iy.numpy().astype(np.int32)
    # FMA vs non-FMA: the C oracle in CPU (non-FMA) mode must equal the torch restatement bit for bit
    oc, odx, oix, ody, oiy = orc.chamfer_fwd(x.numpy(), y.numpy(), use_fma=False)
    tdx, _ = tr.knn1(x, y)
    assert np.array_equal(oix, ch["c1_idx_x"]) and np.array_equal(oiy, ch["c1_idx_y"])
    assert np.array_equal(odx, tdx.numpy()), "non-FMA oracle distances differ from torch restatement"
    fc, fdx, fix, fdy, fiy = orc.chamfer_fwd(x.numpy(), y.numpy(), use_fma=True)
    ch["c1_fma_idx_x"], ch["c1_fma_idx_y"] = fix, fiy
    ch["c1_fma_cham"] = fc
    print("config1 cham", cham.numpy(), "fma", fc, "idx mismatches fma vs non-fma:",
          int((fix != oix).sum()), int((fiy != oiy).sum()),
          "dist bit diffs:", int((fdx != odx).sum()))

    # ties: cloud tiled x4 (data_loaders.py:80-87) -> lowest index must win
    gg = torch.Generator().manual_seed(5)
    base = torch.rand(1, 300, 3, generator=gg)
    yt = base.repeat(1, 4, 1)
    xt = torch.cat([base[:, :200], torch.rand(1, 311, 3, generator=gg)], dim=1)
    cham_t, ixt, iyt = tr.chamfer_distance(xt, yt)
    ch["tie_x"], ch["tie_y"] = xt.numpy(), yt.numpy()
    ch["tie_cham"], ch["tie_idx_x"], ch["tie_idx_y"] = cham_t.numpy(), ixt.numpy().astype(np.int32), iyt.numpy().astype(np.int32)
    assert int(ixt[0, :200].max()) < 300
    # P1 != P2, B = 3, grads wrt both (autoencoder case: only y)
    xa = torch.rand(3, 700, 3, generator=gg).requires_grad_(True)
    ya = torch.rand(3, 450, 3, generator=gg).requires_grad_(True)
    ca = tr.chamfer_autograd(xa, ya)
    gc = torch.rand(3, generator=gg)
    (ca * gc).sum().backward()
    ch["ne_x"], ch["ne_y"], ch["ne_cham"], ch["ne_gcham"] = xa.detach().numpy(), ya.detach().numpy(), ca.detach().numpy(), gc.numpy()
    ch["ne_gx"], ch["ne_gy"] = xa.grad.numpy(), ya.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "chamfer.npz"), **ch)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KB")


if __name__ == "__main__":
    main()
