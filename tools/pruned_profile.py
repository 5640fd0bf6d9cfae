"""Dev tool (GPU): per-kernel time of one Chamfer forward under an algorithm.  python tools/pruned_profile.py algo kind B P"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import ptk_b200

algo, kind, B, P = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(1)
if kind == "cube":
    x, y = torch.rand(B, P, 3, device=dev, generator=gen) - 0.5, torch.rand(B, P, 3, device=dev, generator=gen) - 0.5
else:
    f = lambda: torch.nn.functional.normalize(torch.randn(B, P, 3, device=dev, generator=gen), dim=-1) * 0.25
    x, y = f() * (1 + 0.02 * torch.randn(B, P, 1, device=dev, generator=gen)), f()
ptk_b200.ops.set_chamfer_algo(algo)
for _ in range(3):
    ptk_b200.ops.chamfer(x, y)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        ptk_b200.ops.chamfer(x, y)
    torch.cuda.synchronize()
print(f"{algo} {kind} B={B} P={P}")
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:8]:
    print(f"  {e.device_time_total / 5:10.1f} us  {e.count // 5}x  {e.key[:90]}")
