"""Dev tool (GPU, library built with PTK_EXTRA_NVCC_FLAGS=-DPTK_PR_STATS): traversal counters of the pruned scan."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ptk_b200

dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(1)
lib = ptk_b200._lib.lib()
buf = (ctypes.c_ulonglong * 16)()
ptk_b200.ops.set_chamfer_algo("pruned")
for kind, B, P in (("cube", 64, 10000), ("cube", 256, 10000), ("sphere", 1, 100000), ("sphere", 8, 50000), ("cube", 64, 2000)):
    if kind == "cube":
        x, y = torch.rand(B, P, 3, device=dev, generator=gen) - 0.5, torch.rand(B, P, 3, device=dev, generator=gen) - 0.5
    else:
        f = lambda: torch.nn.functional.normalize(torch.randn(B, P, 3, device=dev, generator=gen), dim=-1) * 0.25
        x, y = f() * (1 + 0.02 * torch.randn(B, P, 1, device=dev, generator=gen)), f()
    lib.ptk_debug_pr_stats(buf, 1)
    ptk_b200.ops.chamfer(x, y)
    lib.ptk_debug_pr_stats(buf, 1)
    w = max(buf[0], 1)
    print(f"{kind:7s} B={B:3d} P={P:6d}: warps {buf[0]}, per warp: L2 pops {buf[1] / w:.1f}, L1 pops {buf[2] / w:.1f}, leaf tests {buf[3] / w:.1f}, "
          f"leaf scans {buf[4] / w:.1f}, bails {buf[5]}, max per half {buf[6] / w:.1f}, wanting lanes per scan {buf[7] / max(buf[4], 1):.1f}")
    n = 2 * B
    print("         sort kernel, cycles per CTA: " + ", ".join(f"{name} {buf[8 + i] / n:.0f}" for i, name in
                                                              enumerate(("bbox", "hist", "scan", "scatter", "boxes"))))
