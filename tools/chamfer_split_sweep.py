"""Dev tool: forward Chamfer time against the number of target splits (PTK_CH_NSPLIT override) for a few shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ptk_b200
dev = torch.device("cuda")
def timeit(fn, iters):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
shapes = [(10000, 16), (10000, 4), (10000, 1), (10000, 64), (5000, 16), (20000, 16), (50000, 4), (50000, 1), (100000, 1), (4000, 64)]
for P, B in shapes:
    x = torch.rand(B, P, 3, device=dev) - 0.5
    y = torch.rand(B, P, 3, device=dev) - 0.5
    f = lambda: ptk_b200.ops.chamfer(x, y)
    iters = max(3, min(50, int(1e11 / (B * P * P))))
    os.environ.pop("PTK_CH_NSPLIT", None)
    base = timeit(f, iters)
    row = []
    for ns in (1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 14, 16, 20, 24, 30, 40, 60, 80, 120, 160):
        if ns > (P + 255) // 256: break
        os.environ["PTK_CH_NSPLIT"] = str(ns)
        row.append((ns, timeit(f, iters)))
    best = min(row, key=lambda r: r[1])
    print(f"P={P} B={B}: auto {base:.3f} ms; best ns={best[0]} {best[1]:.3f} ms; " + " ".join(f"{n}:{t:.3f}" for n, t in row), flush=True)
