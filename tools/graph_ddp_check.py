"""Dev tool (torchrun, N GPUs): the object-sharded reconstruction step INCLUDING the bucketed NCCL gradient all-reduce
captured in one CUDA graph per rank (recon.GraphedStep) -- parameters after K replays against K eager steps, and the time."""
import sys, os, types, copy, faulthandler
faulthandler.dump_traceback_later(150, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import ptk_b200
from ptk_b200.graph import Graph

rank, world, local = ptk_b200.dist.init_from_env()
dev = torch.device("cuda", local)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
adj = dict(np.load(os.path.join(ROOT, "tests/golden/adjacency.npz")))
meshes = dict(np.load(os.path.join(ROOT, "tests/golden/meshes.npz")))
big = len(sys.argv) > 1 and sys.argv[1] == "big"
args = types.SimpleNamespace(use_img=True, use_touch=True, finger=True, num_grasps=5, num_GCN_layers=20 if big else 6,
                             hidden_GCN_size=300 if big else 120, cut=0.33)
g = Graph.from_csr(adj["p_adj_rowptr"], adj["p_adj_col"], dev)
adj_info = {"origional": Graph.from_csr(adj["p_origional_rowptr"], adj["p_origional_col"], dev).dense(), "adj": g.dense(),
            "faces": torch.from_numpy(adj["p_faces"]).to(dev, torch.int64)}
Bs, width, npts = (16, 448, 10000) if big else (4, 48, 3000)
torch.manual_seed(0)
net0 = ptk_b200.recon.ChartDeformer(adj_info, args, width).to(dev)
gen = torch.Generator(device=dev).manual_seed(100 + rank)
vision = torch.from_numpy(meshes["vision_verts"]).to(dev)[None].repeat(Bs, 1, 1)
touch = torch.rand(Bs, 125, 3, device=dev, generator=gen) * 0.02 + 0.2
feats = [torch.rand(Bs, n, width, device=dev, generator=gen) for n in (1824, 1949, 1949)]
gt = torch.nn.functional.normalize(torch.randn(Bs, npts, 3, device=dev, generator=gen), dim=-1) * 0.25
uni = [ptk_b200.utils.draw_uniforms(Bs, npts, dev, gen) for _ in range(3)]   # fixed draws: eager and graph see the same loss


def make(net):
    opt = torch.optim.Adam(net.parameters(), lr=lr_t, fused=True, capturable=True)
    red = ptk_b200.dist.GradReducer(net.parameters(), bucket_mb=4)

    def step():
        opt.zero_grad(set_to_none=True)
        verts = net(vision, touch, lambda it, v: feats[it])
        _, cd = ptk_b200.recon.recon_loss(verts, adj_info["faces"], gt, number_points=npts, uniforms=uni)
        loss = 9000.0 * cd.sum() / (Bs * world)
        loss.backward()
        red.finish()
        opt.step()
        return loss
    return step


K = 2
lr_t = torch.zeros((), device=dev)   # parity leg: the optimizer runs but leaves the parameters alone, so both arms see the same weights every step
def say(m):
    print(f"[rank {rank}] {m}", flush=True)
say("built")
net_e, net_g = copy.deepcopy(net0), copy.deepcopy(net0)
step_e = make(net_e)
for _ in range(K + 3):          # GraphedStep runs 3 warm-up steps before it captures
    le = step_e()
say("eager done")
graphed = ptk_b200.recon.GraphedStep(make(net_g), warmup=3)
say("captured")
for _ in range(K):
    lg = graphed()
torch.cuda.synchronize()
say("replayed")
err = max(float((a.grad - b.grad).abs().max() / b.grad.abs().max().clamp_min(1e-30)) for a, b in zip(net_g.parameters(), net_e.parameters()))
le, lg = float(le.detach()), float(lg.detach())


def timeit(fn, iters=10):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


lr_t.fill_(3e-4)          # the captured optimizer reads the learning rate from this tensor at replay
te, tg = timeit(step_e), timeit(graphed)
ok = err < 1e-4 and abs(le - lg) < 1e-5 * abs(le)
print(f"rank {rank}/{world}: all-reduced gradients, graph replay vs eager: max rel diff {err:.2e}; loss {lg:.5f} vs {le:.5f}; "
      f"eager {te:.2f} ms, graph replay {tg:.2f} ms -> {'OK' if ok else 'FAIL'}", flush=True)
graphed = None
torch.cuda.synchronize()
if world > 1:
    dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok else 1)
