"""Dev tool: correctness + timing of the GCN linear kernels (SIMT FP32 vs 3xTF32 tcgen05)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ptk_b200
from ptk_b200 import _lib

L = _lib.lib()
dev = torch.device("cuda")

def rel(a, b):
    a = a.double().cpu().numpy(); b = b.double().cpu().numpy()
    return np.abs(a - b).max() / np.abs(b).max()

def timeit(fn, it=20):
    """GPU time per call: the call is captured in a CUDA graph (x10) and replayed, so host overhead
    (python, ctypes, tensor-map encoding) is excluded."""
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
        with torch.cuda.graph(g, stream=s):
            for _ in range(10): fn()
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(max(it // 10, 2)): g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / (max(it // 10, 2) * 10)

shapes = [(128, 32, 16), (128, 64, 160), (256, 300, 300), (1949, 300, 300), (31184, 300, 300), (31184, 448, 300), (29184, 448, 300), (500, 36, 20)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in s.split(",")) for s in sys.argv[1:]]
torch.backends.cuda.matmul.allow_tf32 = False
for (M, K, N) in shapes:
    g = torch.Generator(device=dev).manual_seed(M + K + N)
    X = torch.randn(M, K, device=dev, generator=g)
    W = torch.randn(K, N, device=dev, generator=g) * 0.1
    G = torch.randn(M, N, device=dev, generator=g)
    ref = X.double() @ W.double()
    refd = G.double() @ W.double().t()
    for mode, name in ((1, "simt"), (2, "tf32x3")):
        try:
            H = ptk_b200.ops._linear_fwd(X, W, algo_id=mode)
            torch.cuda.synchronize()
            e = rel(H, ref)
            t = timeit(lambda: ptk_b200.ops._linear_fwd(X, W, algo_id=mode))
            gX = ptk_b200.ops._linear_dgrad(G, W, X, algo_id=mode)
            torch.cuda.synchronize()
            ed = rel(gX, refd * (X > 0))
            td = timeit(lambda: ptk_b200.ops._linear_dgrad(G, W, X, algo_id=mode))
            gW = ptk_b200.ops._linear_wgrad(X, G, algo_id=mode)
            torch.cuda.synchronize()
            ew = rel(gW, X.double().t() @ G.double())
            tw = timeit(lambda: ptk_b200.ops._linear_wgrad(X, G, algo_id=mode))
            print(f"    wgrad err {ew:.2e} {tw*1e3:8.1f} us {2*M*K*N/tw/1e9:8.1f} TFLOPS")
            print(f"M={M} K={K} N={N} {name:7s} fwd err {e:.2e} {t*1e3:8.1f} us {2*M*K*N/t/1e9:8.1f} TFLOPS | dgrad err {ed:.2e} {td*1e3:8.1f} us {2*M*K*N/td/1e9:8.1f} TFLOPS", flush=True)
        except Exception as ex:
            print(f"M={M} K={K} N={N} {name}: {type(ex).__name__}: {ex}", flush=True)
    t = timeit(lambda: torch.matmul(X, W))
    e = rel(torch.matmul(X, W), ref)
    print(f"M={M} K={K} N={N} cublas  fwd err {e:.2e} {t*1e3:8.1f} us {2*M*K*N/t/1e9:8.1f} TFLOPS", flush=True)
