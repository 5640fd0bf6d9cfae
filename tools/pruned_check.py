"""Dev tool (GPU): the pruned Chamfer scan against the filter scan -- keys must be bit-identical on every shape /
distribution -- and the time of both (CUDA events, L2 flushed by the size of the inputs or not at all: small cases
are launch-bound anyway).  python tools/pruned_check.py [quick]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ptk_b200

dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(1)


def clouds(kind, B, P1, P2):
    if kind == "cube":
        return torch.rand(B, P1, 3, device=dev, generator=gen) - 0.5, torch.rand(B, P2, 3, device=dev, generator=gen) - 0.5
    if kind == "sphere":   # surface samples: what the reconstruction loss sees
        f = lambda P: torch.nn.functional.normalize(torch.randn(B, P, 3, device=dev, generator=gen), dim=-1) * 0.25
        return f(P1) * (1 + 0.02 * torch.randn(B, P1, 1, device=dev, generator=gen)), f(P2)
    if kind == "apart":    # disjoint clouds: every bound is loose
        return torch.rand(B, P1, 3, device=dev, generator=gen), torch.rand(B, P2, 3, device=dev, generator=gen) + 3.0
    if kind == "lattice":  # exact ties everywhere
        x = torch.randint(0, 12, (B, P1, 3), device=dev, generator=gen).float() * 0.125
        y = torch.randint(0, 12, (B, P2, 3), device=dev, generator=gen).float() * 0.125
        return x, y
    if kind == "same":     # every point identical
        return torch.full((B, P1, 3), 0.3, device=dev), torch.full((B, P2, 3), 0.3, device=dev)
    if kind == "flat":     # degenerate axis + clusters
        x = torch.rand(B, P1, 3, device=dev, generator=gen); x[..., 2] = 0.5
        y = torch.rand(B, P2, 3, device=dev, generator=gen) * 0.01; y[:, ::2] += 0.9
        return x, y
    if kind == "nan":
        x, y = torch.rand(B, P1, 3, device=dev, generator=gen), torch.rand(B, P2, 3, device=dev, generator=gen)
        x[0, P1 // 2, 1] = float("nan"); y[-1, 0, 0] = float("inf")
        return x, y
    raise ValueError(kind)


def timeit(fn, iters):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
cases = [("cube", 1, 1, 1), ("cube", 2, 15, 17), ("cube", 3, 33, 1000), ("cube", 2, 5000, 7000), ("sphere", 4, 10000, 10000),
         ("apart", 2, 3000, 2000), ("lattice", 2, 4000, 3000), ("same", 2, 600, 500), ("flat", 3, 2500, 4100),
         ("nan", 2, 700, 900), ("cube", 1, 40000, 35000), ("sphere", 1, 100000, 100000)]
if not quick:
    cases += [("cube", 256, 10000, 10000), ("sphere", 256, 10000, 10000), ("cube", 64, 2000, 2000), ("sphere", 16, 50000, 50000),
              ("sphere", 4, 100000, 100000), ("cube", 4, 100000, 100000), ("sphere", 2, 300000, 300000)]
bad = 0
print(f"{'kind':8s} {'B':>4s} {'P1':>7s} {'P2':>7s}  {'filter ms':>10s} {'pruned ms':>10s} {'x':>6s}  rescued   check")
for kind, B, P1, P2 in cases:
    x, y = clouds(kind, B, P1, P2)
    res = {}
    for algo in ("filter", "pruned"):
        ptk_b200.ops.set_chamfer_algo(algo)
        c, ix, iy = ptk_b200.ops.chamfer(x, y)
        d, i = ptk_b200.ops.knn1(x, y)
        n = ptk_b200.ops.chamfer_rescued(x, y)
        iters = 3 if B * P1 * P2 > 1e9 else 10
        ms = timeit(lambda: ptk_b200.ops.chamfer(x, y), iters)
        res[algo] = (c, ix, iy, d, i, n, ms)
    ptk_b200.ops.set_chamfer_algo("filter")
    f, p = res["filter"], res["pruned"]
    nanok = lambda a, b: torch.equal(torch.nan_to_num(a, nan=-1.0), torch.nan_to_num(b, nan=-1.0))
    ok = torch.equal(f[1], p[1]) and torch.equal(f[2], p[2]) and nanok(f[0], p[0]) and nanok(f[3], p[3]) and torch.equal(f[4], p[4])
    bad += not ok
    print(f"{kind:8s} {B:4d} {P1:7d} {P2:7d}  {f[6]:10.3f} {p[6]:10.3f} {f[6] / p[6]:6.2f}  {p[5]:7d}   {'ok' if ok else 'MISMATCH'}", flush=True)
    if not ok:
        for name, a, b_ in (("idx_x", f[1], p[1]), ("idx_y", f[2], p[2]), ("knn idx", f[4], p[4])):
            ne = (a != b_).nonzero()
            if len(ne):
                print(f"   {name}: {len(ne)} differ, first {ne[0].tolist()} filter {a[tuple(ne[0])].item()} pruned {b_[tuple(ne[0])].item()}")
print("FAILED" if bad else "all ok")
sys.exit(1 if bad else 0)
