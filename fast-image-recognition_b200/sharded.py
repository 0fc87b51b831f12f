"""Gallery row-sharding across the GPUs of one box: one process per GPU (torchrun),
rank r holds rows [lo_r, hi_r) of the class-major gallery with index_offset = lo_r, queries are replicated
(a host batch is uploaded in 1/world slices and all-gathered over NVLink, `gather_queries`),
each rank's exact (already reranked) top-k is all-gathered with torch.distributed and merged by (dist, idx).
PNN class scores are summed (all-reduce); per-class minima are min-reduced on packed (dist, idx) keys.

The only data-path collectives are those tiny gathers/reductions (Q*k*8 bytes per rank); the distance work
never leaves its GPU.  `local_search` / `merge` are injectable so the plumbing can be exercised on CPU with
the gloo backend (tests/test_sharded_gloo.py)."""
import numpy as np


def shard_bounds(n, world, rank):
    """Contiguous, balanced row ranges: rank r gets [n*r//world, n*(r+1)//world)."""
    return (n * rank) // world, (n * (rank + 1)) // world


def merge_topk_host(parts_dist, parts_idx, k):
    """numpy reference of the device merge: parts [P, Q, k] → lexicographic (dist, idx) top-k, idx<0 = empty."""
    P, Q, kk = parts_dist.shape
    d = np.transpose(parts_dist, (1, 0, 2)).reshape(Q, P * kk).astype(np.float64)
    i = np.transpose(parts_idx, (1, 0, 2)).reshape(Q, P * kk).astype(np.int64)
    d = np.where(i < 0, np.inf, d)
    order = np.lexsort((i, d), axis=1)[:, :k]
    od = np.take_along_axis(d, order, 1)
    oi = np.take_along_axis(i, order, 1)
    oi = np.where(np.isinf(od), -1, oi)
    return oi.astype(np.int32), np.where(oi < 0, 0, od).astype(np.float32)


class ShardedGallery:
    """Rank-local shard + collectives.  `rows`/`labels` are THIS rank's shard (already sliced with shard_bounds)."""

    def __init__(self, rows, labels, metric, n_total, lo, dist=None, local_factory=None, merge=None, device=None):
        self.dist, self.n_total, self.lo, self.device = dist, int(n_total), int(lo), device
        self.world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
        if local_factory is None:
            import fir_b200
            self.local = fir_b200.Gallery(rows, labels, metric, index_offset=lo)
            if self.world > 1:      # per-class outputs must have the global width on every shard
                import torch
                nc = torch.tensor([self.local.n_classes], dtype=torch.int64, device=rows.device if hasattr(rows, 'device') else 'cpu')
                dist.all_reduce(nc, op=dist.ReduceOp.MAX)
                self.local.set_num_classes(int(nc.item()))
            self._merge = merge or (lambda pd, pi, k: fir_b200.merge_topk(pd, pi))
        else:
            self.local = local_factory(rows, labels, metric, lo)
            self._merge = merge

    def search(self, queries, k=1, **kw):
        idx, dd = self.local.search(queries, k=k, **kw)
        if self.world == 1:
            return idx, dd
        import torch
        t_i = idx if torch.is_tensor(idx) else torch.from_numpy(np.ascontiguousarray(idx))
        t_d = dd if torch.is_tensor(dd) else torch.from_numpy(np.ascontiguousarray(dd))
        nq = t_i.shape[0]
        g_i = torch.empty((self.world * nq, k), dtype=t_i.dtype, device=t_i.device)      # concatenated along dim 0 (gloo and nccl)
        g_d = torch.empty((self.world * nq, k), dtype=t_d.dtype, device=t_d.device)
        self.dist.all_gather_into_tensor(g_i, t_i.contiguous())
        self.dist.all_gather_into_tensor(g_d, t_d.contiguous())
        g_i, g_d = g_i.view(self.world, nq, k), g_d.view(self.world, nq, k)
        if g_d.is_cuda:
            return self._merge(g_d, g_i, k)
        return self._merge(g_d.numpy(), g_i.numpy(), k)

    def gather_queries(self, q_host, device=None):
        """Every rank holds the same query batch in HOST memory (replicated input).  Instead of eight identical H2D copies
        competing for the host's PCIe roots, each rank uploads 1/world of the rows over its own link and the batch is
        assembled on every GPU by an all-gather over NVLink.  Returns the full [nq, d] batch on this rank's device."""
        import torch
        t = q_host if torch.is_tensor(q_host) else torch.from_numpy(np.ascontiguousarray(q_host))
        device = device if device is not None else self.device
        if self.world == 1:
            return t.to(device, non_blocking=True) if device is not None else t
        nq, d = t.shape
        per = -(-nq // self.world)                                    # equal slices (the last one padded)
        rank = self.dist.get_rank()
        lo, hi = min(nq, rank * per), min(nq, (rank + 1) * per)
        part = torch.zeros((per, d), dtype=t.dtype, device=device)
        if hi > lo:
            part[: hi - lo].copy_(t[lo:hi], non_blocking=True)
        full = torch.empty((self.world * per, d), dtype=t.dtype, device=device)
        self.dist.all_gather_into_tensor(full, part)
        return full[:nq]

    def pnn_scores(self, queries, var):
        sc, _ = self.local.pnn_scores(queries, var, n_total=self.n_total)
        import torch
        t = sc if torch.is_tensor(sc) else torch.from_numpy(np.ascontiguousarray(sc))
        if self.world > 1:
            self.dist.all_reduce(t)          # sum of per-shard partial Parzen sums (already divided by n_total)
        lab = torch.argmax(t, dim=1).to(torch.int32)   # first maximum = lowest class on ties (classification.cpp:217-225)
        return (t, lab) if torch.is_tensor(sc) else (t.numpy(), lab.numpy())
