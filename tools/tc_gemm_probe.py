"""Dev tool: GPU time of the tcgen05 wgrad / dgrad / forward kernels at the reconstruction-step shape; PTK_TG_DEBUG
(1 = skip the A split, 2 = skip the drain loads, 4 = skip the MMAs) shows the floors of the dgrad / forward kernel.
(The same switches inside the wgrad kernel cost it 18 us per launch by themselves and were removed again.)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ptk_b200
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
dev = torch.device("cuda")
M, K, N = 31184, 300, 300
X = torch.randn(M, K, device=dev); G = torch.randn(M, N, device=dev); W = torch.randn(K, N, device=dev) * 0.1
def timeit(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
        with torch.cuda.graph(g, stream=s):
            for _ in range(10): fn()
    torch.cuda.synchronize(); g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 50 * 1e3
print(f"TG_DEBUG={os.environ.get('PTK_TG_DEBUG','0')}: "
      f"wgrad {timeit(lambda: ptk_b200.ops._linear_wgrad(X, G, algo_id=2)):.1f} us  "
      f"dgrad {timeit(lambda: ptk_b200.ops._linear_dgrad(G, W, X, algo_id=2)):.1f} us  "
      f"fwd_tc {timeit(lambda: ptk_b200.ops._linear_fwd(X, W, algo_id=2)):.1f} us")
