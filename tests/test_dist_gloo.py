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
    first = [p.grad.clone() for p in net.parameters()]
    # after finish() the gradients ARE the bucket slices (no copy back) ...
    views = {id(p): p.grad.data_ptr() for p in net.parameters()}
    flat_ptrs = {flat[off:off + n].data_ptr() for flat, items in red.buckets for _, off, n in items}
    assert set(views.values()) <= flat_ptrs
    # ... a second step with zero_grad(set_to_none=False) accumulates straight into them
    net.zero_grad(set_to_none=False)
    (net(xs).pow(2).sum(1).sum() / 7).backward()
    red.finish()
    assert all(p.grad.data_ptr() == views[id(p)] for p in net.parameters())
    second = [p.grad.clone() for p in net.parameters()]
    # gradient accumulation: the extra backward inside no_sync(), halves of the shard, same reduced result
    net.zero_grad(set_to_none=True)
    h = (hi - lo) // 2
    with red.no_sync():
        (net(xs[:h]).pow(2).sum(1).sum() / 7).backward()
    (net(xs[h:]).pow(2).sum(1).sum() / 7).backward()
    red.finish()
    third = [p.grad.clone() for p in net.parameters()]
    # two backwards without finish(): refused loudly instead of reducing a bucket that is in flight
    net.zero_grad(set_to_none=True)
    (net(xs).pow(2).sum(1).sum() / 7).backward()
    try:
        (net(xs).pow(2).sum(1).sum() / 7).backward()
        raised = False
    except RuntimeError as e:
        raised = "no_sync" in str(e)
    red.finish()
    # a parameter that gets no gradient on ONE rank only: collectives still pair up (fixed bucket order)
    net.zero_grad(set_to_none=True)
    y = net[0](xs)
    if rank == 0:
        y = net[4](net[3](net[2](net[1](y))))
    (y.pow(2).sum() / 7).backward()
    red.finish()
    lopsided = [p.grad.clone() for p in net.parameters()]
    npy = lambda ts: [t.detach().numpy().copy() for t in ts]  # by value: the worker may exit before the parent reads
    q.put((rank, npy(first), full.numpy().copy(), npy(second), npy(third), raised, npy(lopsided)))
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
    t = lambda arrs: [torch.from_numpy(a) for a in arrs]
    results = [(r, t(g), torch.from_numpy(f), t(s2), t(s3), ra, t(lop)) for r, g, f, s2, s3, ra, lop in results]
    for rank, grads, full, second, third, raised, lopsided in results:
        for name, got in (("first", grads), ("views", second), ("no_sync", third)):
            for g, w in zip(got, want):
                assert torch.allclose(g, w, rtol=1e-5, atol=1e-7), (rank, name)
        assert torch.allclose(full, per_obj.detach(), rtol=1e-6)
        assert raised, "a second backward before finish() must raise"
    # the lopsided step: both ranks hold the same reduced gradients, and the last layer's come from rank 0 alone
    a, b = results[0][6], results[1][6]
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    assert a[-1].abs().max() > 0
