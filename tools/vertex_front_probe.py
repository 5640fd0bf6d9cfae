"""Dev tool (GPU): forward time of the fused vertex front against the CTA count (64 vertices per CTA, 2 CTAs per SM)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ptk_b200
dev = torch.device('cuda')
enc, menc = ptk_b200.Positional_Encoder(448).to(dev), ptk_b200.Mask_Encoder(448).to(dev)
def timeit(fn, iters=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3
print(f"{'vertices':>9} {'CTAs':>5} {'CTAs/SM':>8} {'us':>8}")
for ctas in (148, 296, 444, 457, 488, 520, 592, 740):
    M = ctas * 64
    pos = torch.rand(1, M, 3, device=dev) - 0.5
    mask = torch.randint(0, 4, (1, M, 1), device=dev).float()
    img = torch.rand(1, M, 448, device=dev)
    def fn():
        with torch.no_grad():
            ptk_b200.encoders.vertex_features(enc, menc, pos, mask, img)
    print(f"{M:9d} {ctas:5d} {ctas / 148:8.2f} {timeit(fn):8.1f}")
