"""Dev tool (run under torchrun, N GPUs): sharded candidate scoring equals the single-GPU result."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import ptk_b200

rank, world, local = ptk_b200.dist.init_from_env()
dev = torch.device("cuda", local)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
m = dict(np.load(os.path.join(ROOT, "tests/golden/meshes.npz")))
verts0 = torch.from_numpy(m["obj0_verts"]).to(dev)
faces = torch.from_numpy(m["obj0_faces"].astype(np.int64)).to(dev)
E, A, num = 5, 7, 2000
g = torch.Generator(device="cpu").manual_seed(0)
scale = (1.0 + 0.05 * torch.rand(E, A, 1, 1, generator=g)).to(dev)
verts = verts0[None, None] * scale
gt = (torch.rand(E, 3000, 3, generator=g) * 0.1).to(dev)
uniforms = [(torch.rand(E * A, num, generator=g).to(dev), torch.rand(2, E * A, num, generator=g).to(dev)) for _ in range(3)]
full = ptk_b200.policy.score_candidates(verts, faces, gt, num=num, uniforms=uniforms)
sharded = ptk_b200.policy.score_candidates(verts, faces, gt, num=num, uniforms=uniforms, shard=True)
ok = torch.equal(full, sharded)
a, s = ptk_b200.policy.best_actions(sharded)
print(f"rank {rank}/{world}: sharded scores identical: {ok}; best actions {a.tolist()}", flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok else 1)
