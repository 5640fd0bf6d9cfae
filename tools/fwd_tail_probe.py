"""Dev tool (GPU): time of the exact forward GEMM (K = N = 300) as a function of the number of 64-row tiles -- how much
of the launch at the reconstruction-step shape (M = 31184: 976 CTAs = 2.2 waves of 444) is wave quantisation."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ptk_b200
dev = torch.device("cuda")
K = N = 300
W = torch.randn(K, N, device=dev) * 0.1


def timeit(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


print(f"{'row tiles':>9} {'CTAs':>6} {'waves':>6} {'us':>8} {'us per wave-equivalent':>24}")
for tiles in (148, 222, 296, 370, 444, 456, 488, 518, 592, 666):
    M = tiles * 64
    X = torch.randn(M, K, device=dev)
    us = timeit(lambda: ptk_b200.ops._linear_fwd(X, W, algo_id=1))
    waves = 2 * tiles / 444
    print(f"{tiles:9d} {2 * tiles:6d} {waves:6.2f} {us:8.1f} {us / waves:24.1f}")
