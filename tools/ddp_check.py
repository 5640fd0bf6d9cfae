"""Dev tool (run under torchrun, N GPUs): object-sharded reconstruction step with the bucketed NCCL
gradient all-reduce; checks the reduced gradients against the single-GPU gradient of the full batch."""
import sys, os, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import ptk_b200
from ptk_b200.graph import Graph

rank, world, local = ptk_b200.dist.init_from_env()
dev = torch.device("cuda", local)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
adj = dict(np.load(os.path.join(ROOT, "tests/golden/adjacency.npz")))
meshes = dict(np.load(os.path.join(ROOT, "tests/golden/meshes.npz")))
args = types.SimpleNamespace(use_img=False, use_touch=True, finger=True, num_grasps=5, num_GCN_layers=6, hidden_GCN_size=120, cut=0.33)
g = Graph.from_csr(adj["p_adj_rowptr"], adj["p_adj_col"], dev)
adj_info = {"origional": g.dense(), "adj": g.dense(), "faces": torch.from_numpy(adj["p_faces"]).to(dev, torch.int64)}
Bg = 8
torch.manual_seed(0)
net = ptk_b200.recon.ChartDeformer(adj_info, args, 48).to(dev)
gen = torch.Generator(device="cpu").manual_seed(1)
vision = torch.from_numpy(meshes["vision_verts"])[None].repeat(Bg, 1, 1)
touch = torch.rand(Bg, 125, 3, generator=gen) * 0.02 + 0.2
feats = torch.rand(Bg, 1949, 48, generator=gen)
gt = torch.nn.functional.normalize(torch.randn(Bg, 4000, 3, generator=gen), dim=-1) * 0.25
uni = [(torch.rand(Bg, 4000, generator=gen), torch.rand(2, Bg, 4000, generator=gen)) for _ in range(3)]

def run(lo, hi, reducer=None):
    net.zero_grad(set_to_none=True)
    sl = slice(lo, hi)
    u = [(a[sl].to(dev).contiguous(), b[:, sl].to(dev).contiguous()) for a, b in uni]
    verts = net(vision[sl].to(dev), touch[sl].to(dev), lambda it, v: feats[sl].to(dev))
    _, cd = ptk_b200.recon.recon_loss(verts, adj_info["faces"], gt[sl].to(dev), number_points=4000, uniforms=u)
    (9000.0 * cd.sum() / Bg).backward()
    if reducer is not None:
        reducer.finish()
    return cd.detach(), [p.grad.clone() for p in net.parameters()]

lo, hi = ptk_b200.dist.shard_bounds(Bg, rank, world)
red = ptk_b200.dist.GradReducer(net.parameters(), bucket_mb=0.25)
cd_local, grads = run(lo, hi, red)
cd_all = ptk_b200.dist.gather_objects_vector(cd_local, Bg)
red.enabled = False
cd_ref, grads_ref = run(0, Bg)          # every rank also computes the full batch as ground truth
err = max(float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)) for a, b in zip(grads, grads_ref))
cerr = float((cd_all - cd_ref).abs().max() / cd_ref.abs().max())
ok = err < 1e-5 and cerr < 1e-6
print(f"rank {rank}/{world}: buckets={len(red.buckets)} max grad rel err {err:.2e}, loss gather err {cerr:.2e} -> {'OK' if ok else 'FAIL'}", flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok else 1)
