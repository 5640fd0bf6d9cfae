"""Dev tool (GPU): per-kernel time of one fused vertex-front forward + backward at the config-3 shape."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ptk_b200
from torch.profiler import profile, ProfilerActivity
dev = torch.device('cuda')
enc, menc = ptk_b200.Positional_Encoder(448).to(dev), ptk_b200.Mask_Encoder(448).to(dev)
Bs = 16
pos = (torch.rand(Bs, 1949, 3, device=dev) - 0.5).requires_grad_(True)
mask = torch.randint(0, 4, (Bs, 1949, 1), device=dev).float()
img = torch.rand(Bs, 1949, 448, device=dev); w = torch.rand(Bs, 1949, 448, device=dev)
def fn():
    pos.grad = None; enc.zero_grad(set_to_none=True); menc.zero_grad(set_to_none=True)
    (ptk_b200.encoders.vertex_features(enc, menc, pos, mask, img) * w).sum().backward()
for _ in range(3): fn()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5): fn()
    torch.cuda.synchronize()
tot = 0
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:16]:
    print(f"  {e.device_time_total / 5:9.1f} us  {e.count // 5}x  {e.key[:100]}")
    tot += e.device_time_total / 5
print(f"  total {tot:.1f} us")
