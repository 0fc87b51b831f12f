"""Gallery row-sharding across the GPUs of one box.

The product path lives in the library: csrc/sharded.cu (`fir_comm_*`, `fir_shard_*`, `fir_sharded_*` in
include/fir_b200.h; Python faces `fir_b200.Comm`, `fir_b200.RankShard`, `fir_b200.Sharded`) — NCCL all-gather of packed
64-bit (ordered distance, global index) keys + k-way merge, all-reduce(min) of the same keys for per-class minima,
all-reduce(sum) of the fp64 Parzen partial sums.

This module holds what is independent of the device: the shard bounds, the key format and a host model of the exchange
(`ShardedGallery`) over any `torch.distributed` backend with an injectable rank-local matcher, which is how the protocol
is exercised on CPU with gloo at world_size 2 (tests/test_sharded_gloo.py): same keys, same collectives, same merge rule
as the CUDA kernels (pack_topk_keys_kernel / merge_keys_kernel / unpack_keys_kernel)."""
import numpy as np

EMPTY_KEY = np.uint64(0xFFFFFFFFFFFFFFFF)


def shard_bounds(n, world, rank):
    """Contiguous, balanced row ranges: rank r gets [n*r//world, n*(r+1)//world) — fir_sharded_create uses the same cut."""
    return (n * rank) // world, (n * (rank + 1)) // world


def ordered_bits(dist):
    """fp32 → uint32 whose unsigned order is the float order (fir_common.cuh: ordered_bits)."""
    b = np.ascontiguousarray(dist, dtype=np.float32).view(np.uint32)
    return np.where(b & np.uint32(0x80000000), ~b, b | np.uint32(0x80000000)).astype(np.uint32)


def from_ordered_bits(o):
    o = np.ascontiguousarray(o, dtype=np.uint32)
    b = np.where(o & np.uint32(0x80000000), o & np.uint32(0x7FFFFFFF), ~o).astype(np.uint32)
    return b.view(np.float32)


def pack_keys(dist, idx):
    """(dist, global idx) → uint64 keys, idx < 0 → EMPTY_KEY; key order = lexicographic (dist, idx)."""
    idx = np.asarray(idx)
    keys = (ordered_bits(dist).astype(np.uint64) << np.uint64(32)) | (idx.astype(np.int64) & 0xFFFFFFFF).astype(np.uint64)
    return np.where(idx < 0, EMPTY_KEY, keys)


def unpack_keys(keys, empty_dist=0.0):
    keys = np.asarray(keys, dtype=np.uint64)
    empty = keys == EMPTY_KEY
    d = from_ordered_bits((keys >> np.uint64(32)).astype(np.uint32))
    i = (keys & np.uint64(0xFFFFFFFF)).astype(np.uint32).astype(np.int64)
    return np.where(empty, np.float32(empty_dist), d).astype(np.float32), np.where(empty, -1, i).astype(np.int32)


def merge_keys_host(gathered, k):
    """[world, Q, k] sorted key lists → the k smallest keys per query (what merge_keys_kernel computes)."""
    W, Q, kk = gathered.shape
    flat = np.transpose(gathered, (1, 0, 2)).reshape(Q, W * kk)
    return np.sort(flat, axis=1)[:, :k]


def merge_topk_host(parts_dist, parts_idx, k):
    """numpy reference of the device merge: parts [P, Q, k] → lexicographic (dist, idx) top-k, idx<0 = empty."""
    d, i = unpack_keys(merge_keys_host(pack_keys(parts_dist, parts_idx), k))
    return i, d


def _to_signed(keys):          # order-preserving uint64 → int64 (torch / gloo reduce signed integers)
    return (np.asarray(keys, dtype=np.uint64) ^ np.uint64(1 << 63)).view(np.int64)


def _to_unsigned(s):
    return np.asarray(s, dtype=np.int64).view(np.uint64) ^ np.uint64(1 << 63)


class ShardedGallery:
    """Host model of one rank: `local` answers for THIS rank's rows (global indices), the collectives run over `dist`.
    local_factory(rows, labels, metric, lo) must return an object with search(q, k) → (idx, dist), optionally
    class_min(q) → (min, arg) and pnn_scores(q, var, n_total) → (scores, labels), on numpy arrays."""

    def __init__(self, rows, labels, metric, n_total, lo, dist=None, local_factory=None, merge=None, device=None):
        self.dist, self.n_total, self.lo, self.device = dist, int(n_total), int(lo), device
        self.world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
        if local_factory is None:
            raise ValueError("the CUDA path is fir_b200.RankShard / fir_b200.Sharded (csrc/sharded.cu); this class is the host model "
                             "of the exchange and needs a rank-local matcher")
        self.local = local_factory(rows, labels, metric, lo)

    def _torch(self, a):
        import torch
        t = torch.from_numpy(np.ascontiguousarray(a))
        return t.to(self.device) if self.device is not None else t

    def search(self, queries, k=1, **kw):
        idx, dd = self.local.search(queries, k=k, **kw)
        keys = pack_keys(dd, idx)                                      # one 64-bit key per (query, rank): ONE all-gather
        if self.world == 1:
            out = keys
        else:
            import torch
            t = self._torch(_to_signed(keys))
            g = torch.empty((self.world * t.shape[0], k), dtype=t.dtype, device=t.device)
            self.dist.all_gather_into_tensor(g, t.contiguous())
            out = merge_keys_host(_to_unsigned(g.cpu().numpy()).reshape(self.world, t.shape[0], k), k)
        d, i = unpack_keys(out)
        return i, d

    def class_min(self, queries):
        """Per-class nearest neighbour over all shards: all-reduce(min) of packed (dist, idx) keys; (100000, -1) = none."""
        mn, arg = self.local.class_min(queries)
        keys = pack_keys(mn, arg)
        if self.world > 1:
            t = self._torch(_to_signed(keys))
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
            keys = _to_unsigned(t.cpu().numpy())
        return unpack_keys(keys, empty_dist=100000.0)

    def gather_queries(self, q_host, device=None):
        """Every rank holds the same query batch in HOST memory.  Each rank uploads 1/world of the rows and the batch is
        assembled everywhere by an all-gather (shard_prepare + shard_gather_queries in csrc/sharded.cu)."""
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
        t = self._torch(np.ascontiguousarray(sc, dtype=np.float64))
        if self.world > 1:
            self.dist.all_reduce(t)          # sum of per-shard partial Parzen sums (already divided by n_total)
        s = t.cpu().numpy()
        return s, np.argmax(s, axis=1).astype(np.int32)   # first maximum = lowest class on ties (classification.cpp:217-225)
