"""Dev tool: a few launches of the exact forward GEMM at the reconstruction-step shape (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ptk_b200
dev = torch.device("cuda")
M, K, N = 31184, 300, 300
X = torch.randn(M, K, device=dev); W = torch.randn(K, N, device=dev) * 0.1
for _ in range(6):
    H = ptk_b200.ops._linear_fwd(X, W, algo_id=1)
torch.cuda.synchronize()
