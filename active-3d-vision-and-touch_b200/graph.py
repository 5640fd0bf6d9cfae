"""Sparse form of the GCN adjacency.

The reference keeps the adjacency as a dense, row-normalised (Nv,Nv) fp32 tensor
(pterotactyl/utility/utils.py:47-71) and multiplies it densely
(pterotactyl/reconstruction/vision/model.py:356,360) although it is 99 % zeros.  Callers keep
passing that dense tensor (the GCN_layer.forward signature is unchanged); `graph_of` derives the CSR
once per tensor and caches it.
"""
import weakref

import numpy as np
import torch

HUB_DEG = 128  # rows with more neighbours are handled by a whole CTA (csrc/gcn_aggregate.cu)
COMMON_MIN = 64  # smallest shared neighbour set worth factoring out of the hub rows


def factor_hubs(rowptr, col, val, n):
    """Kernel-side form of one CSR (see ptk_gcn_aggregate_ex in include/ptk.h).

    The touch-chart centre vertices are all linked to the same ~1150 boundary vertices
    (utils.py:126-128).  If the hub rows (degree > HUB_DEG) share a column set S on which
    val[h, j] == alpha[h] * w[j] (row-normalised A^: alpha = 1/deg_h, w = 1; its transpose: alpha = 1,
    w = 1/deg_j), S is split off: the kernel sums it once per batch element.  Returns a dict with the
    (possibly reduced) CSR, the hub list and -- when factored -- common_col / common_w / alpha / row_skip.
    """
    deg = np.diff(rowptr)
    hubs = np.nonzero(deg > HUB_DEG)[0].astype(np.int32)
    plain = dict(rowptr=rowptr, col=col, val=val, hubs=hubs, common_col=None, common_w=None, alpha=None,
                 row_skip=None)
    if len(hubs) < 2:
        return plain
    rows = [(col[rowptr[h]:rowptr[h + 1]], val[rowptr[h]:rowptr[h + 1]]) for h in hubs]
    common = rows[0][0]
    for c, _ in rows[1:]:
        common = np.intersect1d(common, c, assume_unique=True)
    common = np.setdiff1d(common, hubs, assume_unique=True)  # keep hub-hub links in the rows themselves
    if len(common) < COMMON_MIN:
        return plain
    sub = np.stack([v[np.searchsorted(c, common)] for c, v in rows]).astype(np.float64)  # (n_hubs, |S|)
    if (sub == 0).any():
        return plain
    # rank-1 test: prefer an exact form (all rows constant -> w = 1; all rows equal -> alpha = 1)
    if (sub == sub[:, :1]).all():
        alpha, w = sub[:, 0], np.ones(len(common))
    elif (sub == sub[:1]).all():
        alpha, w = np.ones(len(hubs)), sub[0]
    else:
        w = sub[0]
        alpha = sub[:, 0] / w[0]
        if np.abs(np.outer(alpha, w) - sub).max() > 1e-7 * np.abs(sub).max():
            return plain
    keep = np.ones(len(col), bool)
    for h, (c, _) in zip(hubs, rows):
        keep[rowptr[h] + np.searchsorted(c, common)] = False
    ndeg = deg.copy()
    ndeg[hubs] -= len(common)
    nrowptr = np.zeros(n + 1, np.int32)
    np.cumsum(ndeg, out=nrowptr[1:])
    row_skip = np.zeros(n, np.uint8)
    row_skip[hubs] = 1
    return dict(rowptr=nrowptr, col=np.ascontiguousarray(col[keep]), val=np.ascontiguousarray(val[keep]), hubs=hubs,
                common_col=common.astype(np.int32), common_w=w.astype(np.float32), alpha=alpha.astype(np.float32),
                row_skip=row_skip)


TILE_ROWS = 8  # rows per tile of the shared-memory union kernel (csrc/gcn_aggregate_union.cu: AU_TV)


def tile_unions(rowptr, col, skip, n, tile_rows=TILE_ROWS):
    """Per tile of `tile_rows` consecutive rows: the sorted union of the neighbour columns of its rows (rows flagged in
    `skip` -- the hub rows, which dedicated CTAs process -- left out), and per CSR entry the position of its column in
    its tile's union (ptk_gcn_aggregate_tiled: the dense-tile form reads every union row once per batch element, the
    ring form stages the rows in shared memory).  Returns (uptr (n_tiles+1) i32, ucol i32, lidx (nnz) u16, largest union)."""
    n_tiles = (n + tile_rows - 1) // tile_rows
    uptr = np.zeros(n_tiles + 1, np.int32)
    ucols = []
    lidx = np.zeros(len(col), np.uint16)
    max_union = 0
    for t in range(n_tiles):
        r0, r1 = t * tile_rows, min(n, (t + 1) * tile_rows)
        rows = [i for i in range(r0, r1) if not skip[i]]
        if rows:
            ent = np.concatenate([col[rowptr[i]:rowptr[i + 1]] for i in rows])
            u = np.unique(ent)
            for i in rows:
                lidx[rowptr[i]:rowptr[i + 1]] = np.searchsorted(u, col[rowptr[i]:rowptr[i + 1]]).astype(np.uint16)
        else:
            u = np.zeros(0, np.int32)
        ucols.append(u.astype(np.int32))
        uptr[t + 1] = uptr[t] + len(u)
        max_union = max(max_union, len(u))
    ucol = np.concatenate(ucols) if ucols else np.zeros(0, np.int32)
    if max_union > 65535:
        return None
    return uptr, (ucol if len(ucol) else np.zeros(1, np.int32)), (lidx if len(lidx) else np.zeros(1, np.uint16)), max_union


class KernelCSR:
    """Device arrays of one direction (A^ or its transpose) in the form the aggregate kernels take."""

    def __init__(self, f, device):
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device) if a is not None and len(a) else None
        self.rowptr, self.col, self.val = to(f["rowptr"]), to(f["col"]), to(f["val"])
        self.hubs = to(f["hubs"])
        self.n_hubs = int(len(f["hubs"]))
        self.common_col, self.common_w, self.alpha, self.row_skip = (to(f[k]) for k in
                                                                      ("common_col", "common_w", "alpha", "row_skip"))
        self.n_common = 0 if f["common_col"] is None else int(len(f["common_col"]))
        # tile unions of this form (hub rows: the flagged rows of the factored form, else every row above HUB_DEG)
        n = len(f["rowptr"]) - 1
        skip = f["row_skip"] if f["row_skip"] is not None else (
            (np.diff(f["rowptr"]) > HUB_DEG) if len(f["hubs"]) else np.zeros(n, bool))
        self.tile_uptr = self.tile_ucol = self.tile_lidx = None
        self.max_union = 0
        if n > 0 and len(f["col"]) and int(np.diff(f["rowptr"])[~np.asarray(skip, bool)].max(initial=0)) <= HUB_DEG:
            tu = tile_unions(f["rowptr"], f["col"], skip, n)
            if tu is not None:
                uptr, ucol, lidx, self.max_union = tu
                self.tile_uptr = torch.from_numpy(uptr).to(device)
                self.tile_ucol = torch.from_numpy(ucol).to(device)
                # (torch has no uint16 arithmetic, the tensor is only a device buffer: view as int16)
                self.tile_lidx = torch.from_numpy(lidx.view(np.int16)).to(device)


class Graph:
    """CSR of A^ (forward gather) and of A^T (backward gather), resident on one device."""

    def __init__(self, rowptr, col, val, n, device, sym_val_t=None):
        self.n = int(n)
        self.nnz = int(len(col))
        self.device = torch.device(device)
        rowptr = np.ascontiguousarray(rowptr, np.int32)
        col = np.ascontiguousarray(col, np.int32)
        val = np.ascontiguousarray(val, np.float32)
        if sym_val_t is not None:
            # symmetric pattern (device builder): the transpose shares rowptr / col, only the values differ
            rowptr_t, col_t, val_t = rowptr, col, np.ascontiguousarray(sym_val_t, np.float32)
        else:
            # transpose: (i, j, v) -> row j
            rows = np.repeat(np.arange(self.n, dtype=np.int32), np.diff(rowptr))
            order = np.lexsort((rows, col))
            col_t = rows[order]
            val_t = val[order]
            rowptr_t = np.zeros(self.n + 1, np.int32)
            np.cumsum(np.bincount(col, minlength=self.n), out=rowptr_t[1:])
        self.host = dict(rowptr=rowptr, col=col, val=val, rowptr_t=rowptr_t, col_t=col_t, val_t=val_t)
        hubs = np.nonzero(np.diff(rowptr) > HUB_DEG)[0].astype(np.int32)
        hubs_t = np.nonzero(np.diff(rowptr_t) > HUB_DEG)[0].astype(np.int32)
        to = lambda a: torch.from_numpy(a).to(self.device)
        self.rowptr, self.col, self.val = to(rowptr), to(col), to(val)
        self.rowptr_t, self.col_t, self.val_t = to(rowptr_t), to(col_t), to(val_t)
        self.hubs = to(hubs) if len(hubs) else None
        self.hubs_t = to(hubs_t) if len(hubs_t) else None
        self.n_hubs, self.n_hubs_t = int(len(hubs)), int(len(hubs_t))
        # kernel-side forms with the hub rows' shared neighbour set factored out (vector path only)
        self.fwd_k = KernelCSR(factor_hubs(rowptr, col, val, self.n), self.device)
        self.bwd_k = KernelCSR(factor_hubs(rowptr_t, col_t, val_t, self.n), self.device)

    def csr_struct(self, transpose=False):
        """ctypes ptk_gcn_csr (include/ptk.h) of A^ or its transpose, built once; the device arrays it points at are
        owned by this Graph."""
        import ctypes as C

        from . import _lib
        cache = self.__dict__.setdefault("_csr_structs", {})
        key = bool(transpose)
        hit = cache.get(key)
        if hit is None:
            if transpose:
                rp, col, val, hubs, nh, k = self.rowptr_t, self.col_t, self.val_t, self.hubs_t, self.n_hubs_t, self.bwd_k
            else:
                rp, col, val, hubs, nh, k = self.rowptr, self.col, self.val, self.hubs, self.n_hubs, self.fwd_k
            ptr = lambda t: None if t is None else t.data_ptr()
            hit = cache[key] = _lib.GcnCsr(
                ptr(rp), ptr(col), ptr(val), ptr(hubs), nh, ptr(k.rowptr), ptr(k.col), ptr(k.val), ptr(k.hubs), k.n_hubs,
                ptr(k.common_col), ptr(k.common_w), k.n_common, ptr(k.alpha), ptr(k.row_skip),
                ptr(k.tile_uptr), ptr(k.tile_ucol), ptr(k.tile_lidx), k.max_union)
        return hit

    @staticmethod
    def from_dense(adj):
        a = adj.detach().to("cpu", torch.float32).numpy()
        n = a.shape[0]
        if a.ndim != 2 or a.shape[1] != n:
            raise ValueError(f"adjacency must be square, got {tuple(a.shape)}")
        r, c = np.nonzero(a)
        rowptr = np.zeros(n + 1, np.int32)
        np.cumsum(np.bincount(r, minlength=n), out=rowptr[1:])
        return Graph(rowptr, c.astype(np.int32), a[r, c], n, adj.device)

    @staticmethod
    def from_csr(rowptr, col, device, val=None):
        """Row-normalised graph (every entry of row i is fl32(1/deg_i), utils.py:47-52)."""
        rowptr = np.asarray(rowptr, np.int32)
        deg = np.diff(rowptr)
        if val is None:
            with np.errstate(divide="ignore"):
                w = np.where(deg > 0, np.float32(1.0) / deg.astype(np.float32), np.float32(0.0))
            val = np.repeat(w.astype(np.float32), deg)
        return Graph(rowptr, col, val, len(rowptr) - 1, device)

    @staticmethod
    def from_faces(faces, n, positions=None, centres=None):
        """Row-normalised adjacency straight from faces, built on the device (csrc/adjacency.cu; replaces
        calc_adj + the fusing loops of adj_fuse_touch + normalize_adj, utils.py:47-148).

        faces (F,3) integer CUDA tensor with ids in [0,n); positions (n_pos,3) f32: vertices with byte-identical
        positions are linked to each other and to every id in `centres` (both ways)."""
        from . import _lib
        if not faces.is_cuda:
            raise RuntimeError("Graph.from_faces builds on the GPU: faces must be a CUDA tensor (no CPU fallback)")
        dev = faces.device
        if faces.dim() != 2 or faces.shape[1] != 3:
            raise ValueError(f"faces must be (F,3), got {tuple(faces.shape)}")
        n = int(n)
        F = int(faces.shape[0])
        if F and (int(faces.min()) < 0 or int(faces.max()) >= n):
            raise ValueError(f"face vertex ids must lie in [0, {n})")
        f32 = faces.to(torch.int32).contiguous()
        pos = cen = None
        n_pos = n_cen = 0
        if positions is not None and positions.shape[0] > 1:
            if positions.dim() != 2 or positions.shape[1] != 3 or positions.shape[0] > n:
                raise ValueError(f"positions must be (n_pos <= {n}, 3), got {tuple(positions.shape)}")
            pos = positions.detach().to(dev, torch.float32).contiguous()
            n_pos = int(pos.shape[0])
        if centres is not None and len(centres):
            cen = torch.as_tensor(centres, dtype=torch.int32).to(dev).contiguous()
            n_cen = int(cen.numel())
            if int(cen.min()) < 0 or int(cen.max()) >= n:
                raise ValueError(f"centre ids must lie in [0, {n})")
        L = _lib.lib()
        ws_bytes = L.ptk_adj_workspace_bytes(n)
        if ws_bytes == 0:
            raise ValueError(f"graph of {n} vertices is outside the device builder's range")
        ptr = lambda t: 0 if t is None else t.data_ptr()
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream().cuda_stream
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            rowptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
            _lib.check(L.ptk_adj_count(ptr(f32), F, n, ptr(pos), n_pos, ptr(cen), n_cen, ptr(rowptr), ptr(ws), ws_bytes,
                                       stream), "ptk_adj_count")
            nnz = int(rowptr[-1])  # the one host read of the build: sizes col / val
            col = torch.empty(nnz, dtype=torch.int32, device=dev)
            val = torch.empty(nnz, dtype=torch.float32, device=dev)
            val_t = torch.empty(nnz, dtype=torch.float32, device=dev)
            _lib.check(L.ptk_adj_emit(n, ptr(rowptr), ptr(col), ptr(val), ptr(val_t), ptr(ws), ws_bytes, stream),
                       "ptk_adj_emit")
        # ~100-250 KB back to the host for the hub factoring (factor_hubs); the dense matrix never exists
        return Graph(rowptr.cpu().numpy(), col.cpu().numpy(), val.cpu().numpy(), n, dev, sym_val_t=val_t.cpu().numpy())

    def dense(self):
        """The reference's dense row-normalised (n,n) tensor (only for callers that insist on it)."""
        a = torch.zeros(self.n, self.n, device=self.device)
        rows = torch.repeat_interleave(torch.arange(self.n, device=self.device), torch.diff(self.rowptr.long()))
        a[rows, self.col.long()] = self.val
        return a


_cache = {}


def graph_of(adj):
    """Graph for a dense adjacency tensor, cached on (storage, version, shape, device)."""
    if isinstance(adj, Graph):
        return adj
    key = (adj.data_ptr(), adj._version, tuple(adj.shape), str(adj.device))
    hit = _cache.get(key)
    if hit is not None and hit[0]() is not None:
        return hit[1]
    g = Graph.from_dense(adj)
    try:
        ref = weakref.ref(adj)
    except TypeError:
        ref = lambda: adj
    if len(_cache) > 64:
        _cache.clear()
    _cache[key] = (ref, g)
    return g


def register(adj, graph):
    """Pre-populate the cache (adj_init does this so the dense tensor is never re-scanned)."""
    key = (adj.data_ptr(), adj._version, tuple(adj.shape), str(adj.device))
    _cache[key] = (weakref.ref(adj), graph)
    return graph
