"""Dev tool: bit-level checksum of the exact forward GEMM (plain and fused-layer form) for A/B runs of kernel variants
(PTK_FWD_NO_TMA=1 selects the register-staged kernel): the printed checksums must be identical."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ptk_b200
from ptk_b200 import _lib
dev = torch.device("cuda")
for (M, K, N) in [(31184, 300, 300), (29184, 300, 300), (31184, 448, 300), (1949, 300, 300), (3898, 300, 300), (9500, 448, 300),
                  (100, 52, 64), (64, 16, 160), (63, 20, 68)]:
    g = torch.Generator(device=dev).manual_seed(M + K + N)
    X = torch.randn(M, K, device=dev, generator=g)
    W = torch.randn(K, N, device=dev, generator=g) * 0.1
    H = ptk_b200.ops._linear_fwd(X, W, algo_id=1)
    torch.cuda.synchronize()
    ref = (X.double() @ W.double())
    err = float((H.double() - ref).abs().max() / ref.abs().max())
    print(f"M={M} K={K} N={N} checksum {int(H.view(torch.int32).to(torch.int64).sum())} err {err:.2e}")
