"""Dev tool: BASELINE configs[4] sweep -- Chamfer fwd and fwd+bwd over cloud sizes and batch sizes on one GPU.
    python tools/chamfer_sweep.py > profiles/rNN_chamfer_sweep.txt
Each cell also checks size-independent properties (reported index reproduces the reported distance; symmetry)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ptk_b200

dev = torch.device("cuda")
ptk_b200.ops.set_chamfer_algo("filter")   # the main columns are the brute-force scan at every size; the last two the pruned scan

def timeit(fn, iters):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters

def graph_time(fn, iters):
    """GPU time of fn replayed from a CUDA graph: what the kernels cost without the eager host path (autograd
    engine, ctypes, allocator) -- every ptk_b200 op is capture-safe."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn(); fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return timeit(g.replay, iters)

print(f"{'P':>7} {'B':>4} {'fwd ms':>10} {'fwd+bwd ms':>11} {'graph ms':>9} {'pairs/s':>10} {'Tevals/s':>9} {'rescued %':>9} "
      f"{'pruned fwd':>10} {'x':>6}  checks")
for P in (1000, 2000, 5000, 10000, 20000, 50000, 100000):
    for B in (1, 4, 16, 64, 256):
        if B * P * P > 256 * 20000 * 20000 * 1.1:   # keep the sweep within a few seconds per cell
            continue
        g = torch.Generator(device=dev).manual_seed(P + B)
        x = (torch.rand(B, P, 3, device=dev, generator=g) - 0.5).requires_grad_(True)
        y = (torch.rand(B, P, 3, device=dev, generator=g) - 0.5).requires_grad_(True)
        iters = max(2, min(50, int(2e11 / (B * P * P))))
        def fwd():
            with torch.no_grad():
                return ptk_b200.ops.chamfer(x, y)
        def both():
            x.grad = y.grad = None
            c, _, _ = ptk_b200.ops.chamfer(x, y)
            c.sum().backward()
        tf, tb = timeit(fwd, iters), timeit(both, iters)
        tg = graph_time(both, iters) if B * P <= 64 * 10000 else float("nan")
        cham, ix, iy = fwd()
        ar = torch.arange(B, device=dev)[:, None]
        dx = ((x - y[ar, ix.long()]) ** 2).sum(-1).mean(1)
        dy = ((y - x[ar, iy.long()]) ** 2).sum(-1).mean(1)
        ok1 = torch.allclose(cham, (dx + dy).detach(), rtol=1e-5)
        c2, jx, jy = ptk_b200.ops.chamfer(y.detach(), x.detach())
        ok2 = torch.equal(jx, iy) and torch.equal(jy, ix)
        q = x[0, :32].detach()
        ok3 = torch.equal(((q[:, None] - y[0].detach()[None]) ** 2).sum(-1).argmin(1).int(), ix[0, :32])
        resc = ptk_b200.ops.chamfer_rescued(x.detach(), y.detach()) / (2.0 * B * P)
        try:  # the pruned scan on the same clouds: same indices and value, its forward time
            ptk_b200.ops.set_chamfer_algo("pruned")
            tp = timeit(fwd, iters)
            cp, px, py = fwd()
            ok3 = ok3 and torch.equal(px, ix) and torch.equal(py, iy) and torch.equal(cp, cham)
        finally:
            ptk_b200.ops.set_chamfer_algo("filter")
        print(f"{P:7d} {B:4d} {tf:10.3f} {tb:11.3f} {tg:9.3f} {B / (tb * 1e-3):10.1f} {2.0 * B * P * P / (tf * 1e-3) / 1e12:9.3f} "
              f"{100 * resc:9.3f} {tp:10.3f} {tf / tp:6.2f}  {'ok' if (ok1 and ok2 and ok3) else 'FAIL'}", flush=True)
        del x, y
