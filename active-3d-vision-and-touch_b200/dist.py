"""Multi-GPU plumbing: one process per GPU, objects (whole meshes + clouds) sharded over ranks.

The reference is single-process / single-GPU (no distributed code at all).  The hot
path is embarrassingly parallel over the batch dimension -- Chamfer, sampling and the GCN never
exchange data between objects -- so the only collectives are:
  * training: all-reduce (sum) of parameter gradients, bucketed, issued as soon as the backward has
    produced a bucket so that NCCL overlaps the rest of the backward;
  * evaluation / policy scoring: all-gather of the per-object loss / score vector.
Works with backend "nccl" (GPUs, NVLink/NVSwitch) and "gloo" (CPU tests).
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's env (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, world, local


def rank_world():
    """(rank, world) of the default process group, (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n_items, rank, world):
    """Contiguous block [lo, hi) of `n_items` objects owned by `rank` (sizes differ by at most 1)."""
    base, rem = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(t, rank, world, dim=0):
    lo, hi = shard_bounds(t.shape[dim], rank, world)
    return t.narrow(dim, lo, hi - lo)


def gather_objects_vector(local_vec, n_total, group=None):
    """All-gather per-object values (loss / score) into the full (n_total,) vector on every rank."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_vec
    world = dist.get_world_size(group)
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    maxlen = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros(maxlen, dtype=local_vec.dtype, device=local_vec.device)
    pad[: local_vec.numel()] = local_vec
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)])


class GradReducer:
    """Bucketed gradient all-reduce overlapped with the backward.

    Parameters are packed, in reverse registration order (the order the backward produces them), into flat
    buckets of ~bucket_mb.  A post-accumulate-grad hook per parameter copies the gradient into its bucket (no copy
    when `p.grad` already IS the bucket slice, see below); a bucket's all-reduce is launched asynchronously once it
    AND every earlier bucket is complete, so every rank issues the collectives in the same (bucket) order even if a
    parameter receives its gradient late -- or not at all -- on some rank.  `finish()` launches whatever is left,
    waits, and points every `p.grad` at its (reduced) bucket slice: no copy back.  Nothing is divided: the reduction
    is a SUM, so each rank scales its local-sum loss by 1/B_global and the result equals the single-GPU gradient of
    the concatenated batch (vision/train.py:144 `loss.mean()`).

    One backward per `finish()`.  A second backward before `finish()` (gradient accumulation, several losses)
    would overwrite a bucket that may already be in flight: the hook raises instead of reducing garbage.  To
    accumulate over micro-batches, run the extra backwards inside `no_sync()` and only the last one outside.
    With `optimizer.zero_grad(set_to_none=False)` the gradients stay views of the buckets and autograd accumulates
    straight into them.
    """

    def __init__(self, params, bucket_mb=32, group=None, first_bucket_mb=None):
        self.group = group
        self.params = [p for p in params if p.requires_grad]
        self.enabled = dist.is_initialized() and dist.get_world_size(group) > 1
        self.buckets = []  # (flat tensor, [(param, offset, numel)])
        self._pending = {}
        self._handles = []
        self._next = 0  # next bucket to launch: collectives are issued strictly in bucket order
        self._sync = True
        if not self.enabled:
            return
        cap = int(bucket_mb * (1 << 20) / 4)
        cur, cur_n = [], 0
        for p in reversed(self.params):
            if cur and cur_n + p.numel() > cap:
                self._close(cur, cur_n)
                cur, cur_n = [], 0
            cur.append(p)
            cur_n += p.numel()
        if cur:
            self._close(cur, cur_n)
        for bi, (_, items) in enumerate(self.buckets):
            for p, off, n in items:
                p.register_post_accumulate_grad_hook(self._make_hook(bi, off, n))
        self.reset()

    def _close(self, plist, total):
        ref = plist[0]
        flat = torch.zeros(total, dtype=ref.dtype, device=ref.device)
        items, off = [], 0
        for p in plist:
            items.append((p, off, p.numel()))
            off += p.numel()
        self.buckets.append((flat, items))

    def reset(self):
        self._pending = {bi: len(items) for bi, (_, items) in enumerate(self.buckets)}
        self._handles = []
        self._next = 0

    def no_sync(self):
        """Context manager: backwards inside it only accumulate into p.grad (no bucket copy, no collective)."""
        reducer = self

        class _NoSync:
            def __enter__(self):
                reducer._sync = False

            def __exit__(self, *exc):
                reducer._sync = True

        return _NoSync()

    def _launch_ready(self):
        while self._next < len(self.buckets) and self._pending[self._next] == 0:
            flat, _ = self.buckets[self._next]
            self._handles.append((self._next, dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group,
                                                              async_op=True)))
            self._next += 1

    def _make_hook(self, bi, off, n):
        def hook(p):
            if not self._sync:
                return
            if self._pending[bi] <= 0:
                raise RuntimeError(
                    "GradReducer: a parameter received a second gradient before finish() -- its bucket may already "
                    "be in flight. Call finish() after every backward, or run the extra backwards of a gradient "
                    "accumulation inside reducer.no_sync().")
            flat, _ = self.buckets[bi]
            view = flat[off:off + n]
            if p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad.reshape(-1))
            self._pending[bi] -= 1
            self._launch_ready()
        return hook

    def finish(self):
        """Launch the buckets that are still waiting (in order), wait for all of them and make every p.grad the
        reduced slice of its bucket."""
        if not self.enabled:
            return
        for bi in range(self._next, len(self.buckets)):
            if self._pending[bi] > 0:  # parameters that received no gradient this step (on this rank)
                flat, items = self.buckets[bi]
                for p, off, cnt in items:
                    view = flat[off:off + cnt]
                    if p.grad is None:
                        view.zero_()
                    elif p.grad.data_ptr() != view.data_ptr():
                        view.copy_(p.grad.reshape(-1))
                self._pending[bi] = 0
        self._launch_ready()
        for bi, h in self._handles:
            h.wait()
            flat, items = self.buckets[bi]
            for p, off, cnt in items:
                p.grad = flat[off:off + cnt].view_as(p)
        self.reset()
