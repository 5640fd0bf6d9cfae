"""Dev tool: utils.chamfer_distance (fwd + bwd, eager) as one autograd node vs the per-repeat composition, at the shapes of
BASELINE configs[1] (touch charts: B=64, V=25, F=32, 4000 points) and configs[2] (B=16, V=1949, F=2464, 10000 points)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ptk_b200
dev = torch.device("cuda")
def timeit(fn, iters=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
adj = dict(np.load("tests/golden/adjacency.npz")); m = dict(np.load("tests/golden/meshes.npz"))
cases = [("config 2 (touch charts)", torch.from_numpy(m["touch_verts"]).to(dev)[None].repeat(64, 1, 1) + 0.01 * torch.rand(64, 25, 3, device=dev),
          torch.from_numpy(m["touch_faces"].astype(np.int64)).to(dev), 4000, 4000),
         ("config 3 (fused charts)", torch.rand(16, 1949, 3, device=dev) * 0.5, torch.from_numpy(adj["p_faces"]).to(dev, torch.int64), 10000, 10000)]
for name, v0, faces, num, P2 in cases:
    B = v0.shape[0]
    gt = torch.rand(B, P2, 3, device=dev) * 0.5
    for mode in ("uniform", "multinomial"):
        out = []
        for fused in (True, False):
            ptk_b200.utils.fused_mesh_chamfer = fused
            verts = v0.clone().requires_grad_(True)
            def step():
                verts.grad = None
                (9000.0 * ptk_b200.utils.chamfer_distance(verts, faces, gt, num=num, repeat=3, face_draw=mode).mean()).backward()
            n0 = ptk_b200._lib.launch_count(); step(); n1 = ptk_b200._lib.launch_count()
            out.append((timeit(step), n1 - n0))
        ptk_b200.utils.fused_mesh_chamfer = True
        print(f"{name:26s} {mode:12s}: one node {out[0][0]:.3f} ms ({out[0][1]} ptk launches) | per-repeat nodes {out[1][0]:.3f} ms ({out[1][1]} launches)", flush=True)
