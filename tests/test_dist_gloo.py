"""World-size-2 gloo test of the multi-GPU plumbing (CPU): object sharding, bucketed gradient
all-reduce overlapped with backward, per-object loss gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ptk_b200


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model():
    torch.manual_seed(3)
    return torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 16), torch.nn.ReLU(),
                               torch.nn.Linear(16, 3))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = ptk_b200.dist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    net = _model()
    red = ptk_b200.dist.GradReducer(net.parameters(), bucket_mb=0.0005)  # tiny buckets -> several all-reduces
    assert len(red.buckets) > 1
    torch.manual_seed(11)
    X = torch.rand(7, 6)  # 7 objects: uneven shards (4 + 3)
    lo, hi = ptk_b200.dist.shard_bounds(7, rank, world)
    xs = X[lo:hi]
    per_obj = net(xs).pow(2).sum(1)
    (per_obj.sum() / 7).backward()  # global mean = local sum / B_global, reduced with SUM
    red.finish()
    full = ptk_b200.dist.gather_objects_vector(per_obj.detach(), 7)
    q.put((rank, [p.grad.clone() for p in net.parameters()], full))
    dist.barrier()
    dist.destroy_process_group()


def test_grad_allreduce_and_loss_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process ground truth on the concatenated batch
    net = _model()
    torch.manual_seed(11)
    X = torch.rand(7, 6)
    per_obj = net(X).pow(2).sum(1)
    per_obj.mean().backward()
    want = [p.grad for p in net.parameters()]
    for rank, grads, full in results:
        for g, w in zip(grads, want):
            assert torch.allclose(g, w, rtol=1e-5, atol=1e-7), rank
        assert torch.allclose(full, per_obj.detach(), rtol=1e-6)
