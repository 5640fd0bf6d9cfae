timeout 600 python -m pytest tests/test_gcn_gpu.py -x -q -k "forms or aggregate_vs_oracle or random_graphs" 2>&1 | tail -3
python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch, ptk_b200
from ptk_b200.graph import Graph
adj = dict(np.load("tests/golden/adjacency.npz"))
dev = torch.device("cuda")
def timeit(fn, iters, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3
for B, iters in ((256, 20), (16, 100)):
    for name in ("p_adj", "p_origional", "g_adj"):
        g = Graph.from_csr(adj[name + "_rowptr"], adj[name + "_col"], dev)
        for (C, L, ld) in ((300, 99, 0), (300, 300, 0), (100, 99, 300)):
            Hs = [torch.rand(B, g.n, C, device=dev) for _ in range(2)]
            os_ = [torch.empty(B, g.n, ld if ld else C, device=dev) for _ in range(2)]
            bias = torch.rand(C, device=dev)
            res = []
            for form in ("l2", "dense", "ring", "auto"):
                ptk_b200.ops.aggregate_form = form
                k = [0]
                def fn():
                    i = k[0] & 1; k[0] += 1
                    if ld:
                        kk = g.fwd_k; L_ = ptk_b200._lib.lib(); p = ptk_b200.ops._p
                        ptk_b200._lib.check(L_.ptk_gcn_aggregate_tiled(p(kk.rowptr), p(kk.col), p(kk.val), p(kk.hubs), kk.n_hubs, p(kk.common_col), p(kk.common_w), kk.n_common, p(kk.alpha), p(kk.row_skip), p(kk.tile_uptr), p(kk.tile_ucol), p(kk.tile_lidx), kk.max_union, ptk_b200.ops.AGG_FORMS[form], g.n, p(Hs[i]), B, C, L, p(bias), 1, p(os_[i]), C, ld, ptk_b200.ops._stream()), "tiled")
                    else:
                        ptk_b200.ops._aggregate(g, Hs[i], L, bias, True, out=os_[i])
                res.append(timeit(fn, iters))
            ptk_b200.ops.aggregate_form = "auto"
            print(f"B={B:3d} {name:12s} C={C} L={L} ldo={ld or C}: L2 gather {res[0]:7.1f} us | dense tile {res[1]:7.1f} us | ring {res[2]:7.1f} us | auto {res[3]:7.1f} us", flush=True)
PY
