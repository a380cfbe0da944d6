"""Row-sharded database search across GPUs: one process per GPU, the database cut at song
boundaries into contiguous shards, queries replicated.  Per batch of query files:

  1. every rank runs its sample pre-pass and contributes the k best SAMPLED scores per query (one all-gather of
     Q * k * 4 bytes per rank); the filter threshold is the k-th best of the union of all samples -- a lower bound
     of the GLOBAL k-th best score, several times tighter than any single shard's (round 2 first took the maximum
     of the per-shard bounds with an all-reduce: 5x more rows survived the filtered scan at 8 shards);
  2. filtered scan + exact rescoring per shard, then the ONE all-gather of per-shard top-k the north_star asks
     for: [Q, k] sortable 64-bit keys (score bits << 32 | 0xFFFFFFFF - global row id), merged identically on every
     rank into the global top-k (score desc, id asc);
  3. every rank runs the sequence score over the candidates whose songs it owns; one all-gather of a 16-byte
     (score, song, time) record per query file, arg-max with the reference's tie rule (higher score, then lower
     song id, cpp/seqscore.cpp:121) on the device.

Nothing is read back to the host between batches: ``query_batches`` enqueues everything and reads the answers once.
With world size 1 the collectives vanish and the result is exactly ``Database.query_batch``.
The reference has no multi-GPU search (its faiss hook clones replicas, database.py:101-104).
"""
import os

import numpy as np


def shard_songs(song_pos, world):
    """Contiguous song ranges [(s0, s1)] * world with ~equal row counts, cut at song boundaries."""
    song_pos = np.asarray(song_pos, dtype=np.int64)
    n_songs, n = len(song_pos) - 1, int(song_pos[-1])
    cuts = [0]
    for r in range(1, world):
        s = int(np.searchsorted(song_pos, (n * r) // world, side='left'))
        cuts.append(min(max(s, cuts[-1]), n_songs))
    cuts.append(n_songs)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def merge_topk(dists, labels, k):
    """numpy statement of the top-k merge: lists [G][Q][k] -> [Q][k], (valid, score desc, id asc)."""
    D = np.concatenate(list(dists), axis=1)
    I = np.concatenate(list(labels), axis=1)
    invalid = I < 0
    order = np.lexsort((I, -D.astype(np.float64), invalid), axis=1)[:, :k]
    return np.take_along_axis(D, order, 1), np.take_along_axis(I, order, 1)


def pack_keys(D, I):
    """numpy statement of the key packing of pfann_db_search_filtered: order-preserving fp32 bits << 32 |
    (0xFFFFFFFF - id); empty slots (id < 0) -> 0.  Descending key order = score desc, id asc."""
    u = np.ascontiguousarray(D, dtype=np.float32).view(np.uint32).astype(np.uint64)
    flip = np.where(u >> np.uint64(31), u ^ np.uint64(0xFFFFFFFF), u ^ np.uint64(0x80000000))
    key = (flip << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - np.asarray(I).astype(np.uint64) % np.uint64(1 << 32))
    return np.where(np.asarray(I) < 0, np.uint64(0), key)


def unpack_keys(keys):
    keys = np.asarray(keys, dtype=np.uint64)
    hi = (keys >> np.uint64(32)).astype(np.uint32)
    u = np.where(hi >> np.uint32(31), hi ^ np.uint32(0x80000000), hi ^ np.uint32(0xFFFFFFFF)).astype(np.uint32)
    D = u.view(np.float32)
    I = (np.uint64(0xFFFFFFFF) - (keys & np.uint64(0xFFFFFFFF))).astype(np.int64)
    empty = keys == 0
    return np.where(empty, np.float32(-3.4028235e38), D), np.where(empty, -1, I)


def combine_best(scores, songs, times):
    """[G][nq] raw per-shard winners -> global winner per query: score desc, then lower song id; then the
    reference's zero floor (it reads the answer out of the zero-initialised table, database.py:176,190)."""
    scores, songs, times = np.asarray(scores, np.float32), np.asarray(songs, np.int64), np.asarray(times, np.float32)
    nq = scores.shape[1]
    valid = songs >= 0
    has = valid.any(axis=0)
    sc = np.where(valid, scores, -np.inf).astype(np.float64)
    top = sc.max(axis=0)                                            # best score among the shards that have a candidate
    tied = valid & (sc == top[None, :])
    best = np.where(tied, songs, np.iinfo(np.int64).max).argmin(axis=0)   # lowest song id among the tied shards
    cols = np.arange(nq)
    out_g = np.where(has, songs[best, cols], -1).astype(np.int32)
    pos = has & (scores[best, cols] > 0)
    out_s = np.where(pos, scores[best, cols], 0).astype(np.float32)
    out_t = np.where(pos, times[best, cols], 0).astype(np.float32)
    return out_s, out_g, out_t


class ShardedDatabase:
    """Sharded brute-force search + sequence score.  ``backend`` is the per-shard engine; by default the
    libpfann_b200 handle of this rank's shard (``GpuShard(pfann_b200.database.Database(..., songs=range))``).

    Backend contract (tensors live on the backend's device):
      max_norm() / set_max_norm(v)               error bound of the approximate scan, made common to all shards
      thresholds(q, k) -> [Q] fp32                lower bound of the shard's k-th best score per query
      sample_topk(q, k) -> [Q, k] int32 (optional) the k best sampled scores per query as opaque sortable keys, and
      thresholds_from_topk(top_g [G, Q, k], q, k) -> [Q] fp32   lower bound from the union of all shards' samples
      filtered_keys(q, k, thr, defer) -> [Q, k] int64  packed keys of the exact top-k among rows reaching thr
      take_overflow() -> int                       overflowed candidate lists since the last call (deferred mode)
      merge_keys(keys_g [G, Q, k]) -> [Q, k] int64 labels
      rerank_packed(q, query_index, labels, k, fsm, alpha) -> [nq, 4] fp32 (score, song bits, frames, 0)
      combine(packed_g [G, nq, 4]) -> [nq, 4] fp32
    """

    def __init__(self, backend, top_k, frame_shift_mul=1, hop_size=0.5, score_alpha=0.0, group=None):
        self.backend = backend
        self.top_k, self.fsm, self.hop_size, self.alpha = top_k, frame_shift_mul, hop_size, float(score_alpha)
        self.group = group
        import torch
        import torch.distributed as dist
        self.dist = dist if (dist.is_available() and dist.is_initialized()) else None
        self.world = self.dist.get_world_size(group) if self.dist else 1
        if self.world > 1:
            # one error bound for every shard, so that the maximum of the per-shard thresholds stays a lower bound
            m = torch.tensor([backend.max_norm()], dtype=torch.float32, device=backend.torch_device())
            self.dist.all_reduce(m, op=self.dist.ReduceOp.MAX, group=group)
            backend.set_max_norm(float(m.item()))
            scale = os.environ.get('PFANN_B200_SAMPLE_SCALE')
            if hasattr(backend, 'set_sample_scale'):
                # the thresholds come from the union of all shards' samples: each shard can sample less (measured at
                # 2, 4 and 8 shards of a 10 M-row database: 262 k, 98 k and 49 k sampled rows per shard)
                backend.set_sample_scale(float(scale) if scale else (0.67 if self.world < 4 else 0.5))

    def _all_gather(self, t):
        import torch
        if self.world == 1:
            return t.unsqueeze(0)
        out = torch.empty((self.world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        self.dist.all_gather(list(out.unbind(0)), t.contiguous(), group=self.group)   # nccl and gloo alike
        return out

    def search(self, queries, defer=False):
        """Global top-k labels [Q, k] (int64, on the backend's device) of replicated queries."""
        b = self.backend
        q = b.to_device(queries)
        if self.world > 1 and hasattr(b, 'sample_topk'):
            # k-th best of the UNION of the shards' samples: each rank contributes its k best sampled scores per query
            # (Q * k * 4 bytes), every rank reduces the gathered lists itself
            top = b.sample_topk(q, self.top_k)
            thr = b.thresholds_from_topk(self._all_gather(top), q, self.top_k)
        else:
            thr = b.thresholds(q, self.top_k)
            if self.world > 1:
                self.dist.all_reduce(thr, op=self.dist.ReduceOp.MAX, group=self.group)    # Q floats
        keys = b.filtered_keys(q, self.top_k, thr, defer)
        return q, b.merge_keys(self._all_gather(keys), self.top_k)                    # THE all-gather: [G, Q, k] keys

    def _query_dev(self, queries, query_index, defer=False):
        q, labels = self.search(queries, defer)
        packed = self.backend.rerank_packed(q, query_index, labels, self.top_k, self.fsm, self.alpha)
        return self.backend.combine(self._all_gather(packed))                         # [nq, 4], identical on every rank

    def _to_host(self, packed):
        p = packed.cpu().numpy()
        score = np.ascontiguousarray(p[:, 0])
        song = np.ascontiguousarray(p[:, 1]).view(np.int32)
        tim = p[:, 2].astype(np.float64) * self.hop_size / self.fsm                   # database.py:191
        return score, song, tim

    def query_batch(self, queries, query_index):
        """queries [sum len, d], query_index [nq, 2] (start, len), both replicated on every rank.
        Returns (score[nq], song[nq], time_s[nq]) identical on every rank."""
        return self._to_host(self._query_dev(queries, query_index))

    def query_batches(self, queries, query_index, batch):
        """The same for many files, `batch` query files per database pass group; all batches are enqueued before
        the single read-back, so exchanges and small kernels of one batch overlap the host work of the next."""
        import torch
        query_index = np.asarray(query_index, dtype=np.int64).reshape(-1, 2)
        if len(query_index) == 0:
            return np.zeros(0, np.float32), np.zeros(0, np.int32), np.zeros(0, np.float64)

        def run(defer):
            outs = []
            for b0 in range(0, len(query_index), batch):
                qi = query_index[b0:b0 + batch]
                lo, hi = int(qi[:, 0].min()), int((qi[:, 0] + qi[:, 1]).max())
                qi = qi.copy()
                qi[:, 0] -= lo
                outs.append(self._query_dev(queries[lo:hi], qi, defer))
            return self._to_host(torch.cat(outs, dim=0))                              # the one synchronisation

        res = run(True)
        # a candidate list that overflowed somewhere (rare) is only counted in the deferred mode: every rank must
        # take the same decision, then the batches are repeated with the immediate check
        ovf = torch.tensor([self.backend.take_overflow()], dtype=torch.int32, device=self.backend.torch_device())
        if self.world > 1:
            self.dist.all_reduce(ovf, op=self.dist.ReduceOp.MAX, group=self.group)
        return run(False) if int(ovf.item()) else res


class GpuShard:
    """Per-rank engine over libpfann_b200: device-resident, stream-ordered, no host read-backs."""

    def __init__(self, db):
        self.db = db          # pfann_b200.database.Database restricted to this rank's songs

    def torch_device(self):
        import torch
        return torch.device('cuda', self.db.device)

    def to_device(self, queries):
        import torch
        return torch.as_tensor(queries, dtype=torch.float32).to(self.torch_device(), non_blocking=True).contiguous()

    def max_norm(self):
        from . import _lib
        return float(_lib.lib().pfann_db_max_norm(self.db.handle))

    def set_max_norm(self, v):
        from . import _lib
        _lib.check(_lib.lib().pfann_db_set_max_norm(self.db.handle, float(v)), 'pfann_db_set_max_norm')

    def set_sample_scale(self, scale):
        from . import _lib
        _lib.check(_lib.lib().pfann_db_set_sample_scale(self.db.handle, float(scale)), 'pfann_db_set_sample_scale')

    def thresholds(self, q, k):
        import torch
        from . import _lib
        thr = torch.empty(q.shape[0], dtype=torch.float32, device=q.device)
        _lib.use_torch_stream(self.db.device)
        _lib.check(_lib.lib().pfann_db_search_thresholds(self.db.handle, _lib.ptr(q), q.shape[0], k, _lib.ptr(thr)),
                   'pfann_db_search_thresholds')
        return thr

    def sample_topk(self, q, k):
        import torch
        from . import _lib
        top = torch.empty((q.shape[0], k), dtype=torch.int32, device=q.device)
        thr = torch.empty(q.shape[0], dtype=torch.float32, device=q.device)
        _lib.use_torch_stream(self.db.device)
        _lib.check(_lib.lib().pfann_db_search_sample_topk(self.db.handle, _lib.ptr(q), q.shape[0], k, _lib.ptr(thr),
                                                          _lib.ptr(top)), 'pfann_db_search_sample_topk')
        return top

    def thresholds_from_topk(self, top_g, q, k):
        import torch
        from . import _lib
        thr = torch.empty(q.shape[0], dtype=torch.float32, device=q.device)
        _lib.use_torch_stream(self.db.device)
        _lib.check(_lib.lib().pfann_db_thresholds_from_topk(self.db.handle, _lib.ptr(top_g.contiguous()), top_g.shape[0],
                                                            _lib.ptr(q), q.shape[0], k, _lib.ptr(thr)),
                   'pfann_db_thresholds_from_topk')
        return thr

    def filtered_keys(self, q, k, thr, defer=False):
        import torch
        from . import _lib
        keys = torch.empty((q.shape[0], k), dtype=torch.int64, device=q.device)
        _lib.use_torch_stream(self.db.device)
        _lib.check(_lib.lib().pfann_db_search_filtered(self.db.handle, _lib.ptr(q), q.shape[0], k, _lib.ptr(thr),
                                                       _lib.ptr(keys), int(bool(defer))), 'pfann_db_search_filtered')
        return keys

    def take_overflow(self):
        from . import _lib
        return int(_lib.lib().pfann_db_take_overflow(self.db.handle))

    def merge_keys(self, keys_g, k, want_dist=False):
        import torch
        from . import _lib
        G, Q = keys_g.shape[0], keys_g.shape[1]
        labels = torch.empty((Q, k), dtype=torch.int64, device=keys_g.device)
        dist = torch.empty((Q, k), dtype=torch.float32, device=keys_g.device) if want_dist else None
        _lib.use_torch_stream(self.db.device)
        _lib.check(_lib.lib().pfann_topk_merge_keys(_lib.ctx(self.db.device), _lib.ptr(keys_g.contiguous()), G, Q, k,
                                                    _lib.ptr(dist), _lib.ptr(labels)), 'pfann_topk_merge_keys')
        return (dist, labels) if want_dist else labels

    def rerank_packed(self, q, query_index, labels, k, fsm, alpha):
        import torch
        from . import _lib
        qi_host = np.ascontiguousarray(query_index, dtype=np.int64).reshape(-1, 2)
        nq = qi_host.shape[0]
        max_len = int(qi_host[:, 1].max()) if nq else 0
        qi = torch.from_numpy(qi_host).to(q.device, non_blocking=True)
        packed = torch.empty((nq, 4), dtype=torch.float32, device=q.device)
        _lib.use_torch_stream(self.db.device)
        _lib.check(_lib.lib().pfann_db_rerank_packed(self.db.handle, _lib.ptr(q), _lib.ptr(qi), nq, max_len,
                                                     _lib.ptr(labels.contiguous()), k, fsm, float(alpha),
                                                     _lib.ptr(packed)), 'pfann_db_rerank_packed')
        return packed

    def combine(self, packed_g):
        import torch
        from . import _lib
        G, nq = packed_g.shape[0], packed_g.shape[1]
        out = torch.empty((nq, 4), dtype=torch.float32, device=packed_g.device)
        _lib.use_torch_stream(self.db.device)
        _lib.check(_lib.lib().pfann_best_combine(_lib.ctx(self.db.device), _lib.ptr(packed_g.contiguous()), G, nq,
                                                 _lib.ptr(out)), 'pfann_best_combine')
        return out
