"""Dev tool: correctness of the tensor-core forward / dgrad GEMM in the selected PTK_TG_PAIR mode, small to large shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ptk_b200
dev = torch.device("cuda")
shapes = [tuple(int(v) for v in s.split(",")) for s in sys.argv[1:]] or [(256, 64, 160), (128, 64, 160), (1949, 300, 300),
                                                                          (31184, 300, 300), (3000, 448, 300), (777, 300, 448)]
worst = 0.0
for (M, K, N) in shapes:
    g = torch.Generator(device=dev).manual_seed(M + K + N)
    X = torch.randn(M, K, device=dev, generator=g); W = torch.randn(K, N, device=dev, generator=g) * 0.1
    G = torch.randn(M, N, device=dev, generator=g)
    H = ptk_b200.ops._linear_fwd(X, W, algo_id=2); torch.cuda.synchronize()
    ref = X.double() @ W.double()
    e = float((H.double() - ref).abs().max() / ref.abs().max())
    gX = ptk_b200.ops._linear_dgrad(G, W, X, algo_id=2); torch.cuda.synchronize()
    refd = (G.double() @ W.double().t()) * (X > 0)
    ed = float((gX.double() - refd).abs().max() / refd.abs().max())
    print(f"M={M} K={K} N={N}: fwd err {e:.2e} dgrad err {ed:.2e}", flush=True)
    worst = max(worst, e, ed)
sys.exit(0 if worst < 1e-6 else 1)
