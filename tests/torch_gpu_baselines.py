"""Checker-side tool (lives under tests/: it runs the oracle's torch restatement): the bars on the same box -- the reference's own formulation of the GCN part
(dense row-normalised adjacency, torch.matmul: vision/model.py:351-363) and a brute-force torch Chamfer, run with
torch on the GPU (FP32, TF32 off as torch defaults), next to the ptk_b200 kernels on the same inputs.
    python tests/torch_gpu_baselines.py > profiles/rNN_torch_gpu_baselines.txt
PyTorch3D's CUDA knn is not installed here, so the Chamfer bar is torch.cdist + min, not the reference's kernel."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ptk_b200
from ptk_b200.graph import Graph
from oracle import torch_ref as tr

torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
adj = dict(np.load(os.path.join(ROOT, "tests/golden/adjacency.npz")))


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


g = Graph.from_csr(adj["p_adj_rowptr"], adj["p_adj_col"], dev)
dense = g.dense()
print(f"device: {torch.cuda.get_device_name(0)}; graph N={g.n} nnz={g.nnz}")

# 1. aggregation of the 99 propagated channels: dense matmul (reference) vs CSR gather
for B in (16, 256):
    H = torch.rand(B, g.n, 300, device=dev)
    bias = torch.rand(300, device=dev)
    t_ref = timeit(lambda: torch.relu(torch.cat((torch.matmul(dense, H[:, :, :99]) + bias[:99], H[:, :, 99:]), dim=-1)))
    t_ptk = timeit(lambda: ptk_b200.ops._aggregate(g, H, 99, bias, True))
    print(f"aggregate + bias + cat + relu, B={B}: torch dense {t_ref*1e3:8.1f} us | ptk_b200 {t_ptk*1e3:8.1f} us | x{t_ref/t_ptk:.1f}")

# 2. one 20-layer GCN pass (50 -> 300 x 19 -> 3) forward + backward, B = 16
args = types.SimpleNamespace(num_GCN_layers=20, hidden_GCN_size=300, cut=0.33)
torch.manual_seed(0)
net = ptk_b200.GCN(448, args).to(dev)
X = torch.rand(16, g.n, 448, device=dev)
info = {"adj": dense, "origional": dense}
ptk_b200.graph.register(dense, g)
ws, bs = [l.weight for l in net.layers], [l.bias for l in net.layers]


def ref_pass():
    for p in net.parameters():
        p.grad = None
    tr.gcn_dense(X, ws, bs, dense, 0.33).sum().backward()


def ptk_pass():
    for p in net.parameters():
        p.grad = None
    net(X, info).sum().backward()


t_ref, t_ptk = timeit(ref_pass, 3, 1), timeit(ptk_pass, 3, 1)
print(f"GCN pass 448->300x18->3, N=1949, B=16, fwd+bwd: torch dense {t_ref:8.2f} ms | ptk_b200 {t_ptk:8.2f} ms | x{t_ref/t_ptk:.1f}")

# 3. Chamfer 10k x 10k, fwd only, brute force with torch (cdist + min both ways), batches of 4 (1.6 GB of distances)
B = 16
x, y = torch.rand(B, 10000, 3, device=dev) - 0.5, torch.rand(B, 10000, 3, device=dev) - 0.5


def torch_chamfer():
    out = []
    for i in range(0, B, 4):
        d = torch.cdist(x[i:i + 4], y[i:i + 4]).pow(2)
        out.append(d.min(2)[0].mean(1) + d.min(1)[0].mean(1))
    return torch.cat(out)


with torch.no_grad():
    t_ref = timeit(torch_chamfer, 3, 1)
    t_ptk = timeit(lambda: ptk_b200.ops.chamfer(x, y), 10, 2)
    c_ref, c_ptk = torch_chamfer(), ptk_b200.ops.chamfer(x, y)[0]
print(f"Chamfer 10k x 10k fwd, B={B}: torch cdist+min {t_ref:8.2f} ms ({B/t_ref*1e3:7.0f} pairs/s) | ptk_b200 {t_ptk:8.3f} ms "
      f"({B/t_ptk*1e3:7.0f} pairs/s) | x{t_ref/t_ptk:.0f}; values agree to {float(((c_ref-c_ptk).abs()/c_ptk).max()):.1e} "
      f"(cdist's expansion is not the reference's arithmetic)")
