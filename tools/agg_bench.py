"""Dev tool: timing of ptk_gcn_aggregate on the real fused graph at a batch that spills L2.
    python tools/agg_bench.py [B] [iters]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ptk_b200
from ptk_b200.graph import Graph
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
adj = dict(np.load(os.path.join(ROOT, "tests/golden/adjacency.npz")))
dev = torch.device("cuda")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
PEAK = 6554.2

def timeit(fn, iters, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters

for name in ("p_adj", "p_origional", "g_adj"):
    g = Graph.from_csr(adj[name + "_rowptr"], adj[name + "_col"], dev)
    for (C, L, relu, tr) in ((300, 99, True, False), (300, 99, False, True), (300, 300, False, False), (3, 3, False, False)):
        # two buffers of each so that consecutive launches do not hit L2 with the same lines
        Hs = [torch.rand(B, g.n, C, device=dev) for _ in range(2)]
        os_ = [torch.empty(B, g.n, C, device=dev) for _ in range(2)]
        bias = torch.rand(C, device=dev)
        k = [0]
        def fn():
            i = k[0] & 1; k[0] += 1
            ptk_b200.ops._aggregate(g, Hs[i], L, None if tr else bias, relu, transpose=tr, out=os_[i])
        ms = timeit(fn, iters)
        alg = B * g.n * C * 4 * 2 + g.nnz * 8 + (g.n + 1) * 4
        print(f"{name:12s} N={g.n} nnz={g.nnz} B={B} C={C} L={L} relu={int(relu)} T={int(tr)}: {ms*1e3:8.1f} us  "
              f"{alg/ms/1e6:7.1f} GB/s  {alg/ms/1e6/PEAK:.3f} of measured HBM peak", flush=True)
