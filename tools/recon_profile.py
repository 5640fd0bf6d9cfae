"""Dev tool: per-kernel GPU time of one config-3-shaped reconstruction step (torch.profiler)."""
import sys, os, types, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ptk_b200
from ptk_b200.graph import Graph
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
adj = dict(np.load(os.path.join(ROOT, "tests/golden/adjacency.npz")))
meshes = dict(np.load(os.path.join(ROOT, "tests/golden/meshes.npz")))
args = types.SimpleNamespace(use_img=True, use_touch=True, finger=True, num_grasps=5, num_GCN_layers=20, hidden_GCN_size=300, cut=0.33)
g = Graph.from_csr(adj["p_adj_rowptr"], adj["p_adj_col"], dev)
adj_info = {"origional": Graph.from_csr(adj["p_origional_rowptr"], adj["p_origional_col"], dev).dense(), "adj": g.dense(),
            "faces": torch.from_numpy(adj["p_faces"]).to(dev, torch.int64)}
torch.manual_seed(0)
net = ptk_b200.recon.ChartDeformer(adj_info, args, 448).to(dev)
Bs = int(sys.argv[1]) if len(sys.argv) > 1 else 16
vision = torch.from_numpy(meshes["vision_verts"]).to(dev)[None].repeat(Bs, 1, 1)
touch = torch.rand(Bs, 125, 3, device=dev) * 0.02 + 0.2
feats = [torch.rand(Bs, 1824, 448, device=dev), torch.rand(Bs, 1949, 448, device=dev), torch.rand(Bs, 1949, 448, device=dev)]
gt = torch.nn.functional.normalize(torch.randn(Bs, 10000, 3, device=dev), dim=-1) * 0.25
opt = torch.optim.Adam(net.parameters(), lr=3e-4, fused=True)
def step():
    opt.zero_grad(set_to_none=True)
    verts = net(vision, touch, lambda it, v: feats[it])
    loss, _ = ptk_b200.recon.recon_loss(verts, adj_info["faces"], gt, number_points=10000)
    loss.backward()
    opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): step()
b.record(); torch.cuda.synchronize()
print(f"step wall (GPU events): {a.elapsed_time(b)/5:.2f} ms")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
tot = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        tot[e.name[:70]][0] += 1; tot[e.name[:70]][1] += e.device_time
s = sum(v[1] for v in tot.values())
print(f"GPU busy per step: {s/3/1e3:.2f} ms")
for k, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:22]:
    print(f"{t/3/1e3:8.3f} ms {c//3:5d}x {t/c:8.1f} us  {k}")

# ---- idle gaps between consecutive kernels of the last profiled step (where the eager step loses time to the host)
kern = sorted(((e.time_range.start, e.time_range.end, e.name[:48]) for e in prof.events()
               if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda t: t[0])
n3 = len(kern) // 3
kern = kern[2 * n3:]
gaps = []
for (s0, e0, n0), (s1, e1, n1) in zip(kern[:-1], kern[1:]):
    if s1 > e0:
        gaps.append((s1 - e0, n0, n1))
tot_gap = sum(g[0] for g in gaps)
print(f"idle between kernels in one step: {tot_gap/1e3:.2f} ms over {len(gaps)} gaps "
      f"(step span {(kern[-1][1]-kern[0][0])/1e3:.2f} ms, {len(kern)} kernels)")
hist = collections.Counter()
for g, a_, b_ in gaps:
    hist[(a_, b_)] += g
for (a_, b_), g in hist.most_common(14):
    cnt = sum(1 for x in gaps if (x[1], x[2]) == (a_, b_))
    print(f"  {g/1e3:7.3f} ms {cnt:4d}x  {a_}  ->  {b_}")
