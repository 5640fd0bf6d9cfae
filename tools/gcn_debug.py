import sys, os, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ptk_b200
from ptk_b200 import _lib
from ptk_b200.graph import Graph
L = _lib.lib()
adj = dict(np.load("tests/golden/adjacency.npz"))
def rel(a,b): return float((a.double()-b.double()).abs().max()/b.double().abs().max())
torch.manual_seed(1234)
args = types.SimpleNamespace(num_GCN_layers=20, hidden_GCN_size=300, cut=0.33)
net = ptk_b200.GCN(50, args).cuda()
x0 = torch.rand(1, 1949, 50).cuda()
info = {"adj": Graph.from_csr(adj["p_adj_rowptr"], adj["p_adj_col"], "cuda").dense()}
gout = torch.rand(1, 1949, 3, generator=torch.Generator().manual_seed(7)).cuda()
res = {}
for mode in (1 | (1<<2), 1 | (0<<2)):
    L.ptk_gcn_linear_set_mode(mode)
    net.zero_grad()
    x = x0.clone().requires_grad_(True)
    y = net(x, info)
    (y*gout).sum().backward()
    res[mode] = (y.detach().clone(), x.grad.clone(), [l.weight.grad.clone() for l in net.layers], [l.bias.grad.clone() for l in net.layers])
A_,B_ = 1 | (0<<2), 1 | (1<<2)
res = {0: res[A_], 1: res[B_]}
print("y", rel(res[0][0], res[1][0]), "gx", rel(res[0][1], res[1][1]))
for i in range(20):
    print(i, "gW", rel(res[0][2][i], res[1][2][i]), "gb", rel(res[0][3][i], res[1][3][i]))
# single dgrad isolated on realistic data
g = Graph.from_csr(adj["p_adj_rowptr"], adj["p_adj_col"], "cuda")
M=1949
for trial in range(3):
    G = torch.randn(M, 300, device="cuda") * (10.0 ** (-trial*3))
    W = net.layers[5].weight.detach().reshape(300,300)
    A = torch.relu(torch.randn(M,300,device="cuda"))
    outs = {}
    for mode in (1,2):
        L.ptk_gcn_linear_set_mode(mode | (mode<<2))
        outs[mode] = ptk_b200.ops._linear_dgrad(G, W, A)
    ref = (G.double() @ W.double().t()) * (A>0)
    print("dgrad scale", trial, "simt", rel(outs[1], ref), "tc", rel(outs[2], ref))
L.ptk_gcn_linear_set_mode(0)
