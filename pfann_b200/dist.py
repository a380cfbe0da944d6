"""Row-sharded database search across GPUs: one process per GPU, the database cut at song
boundaries into contiguous shards, queries replicated.  Per batch of query files there are exactly two
exchanges, both tiny next to the HBM stream of the scan:

  1. all-gather of the per-shard top-k  [Q, k] x (fp32 score, int64 global row id), merged identically on
     every rank into the global top-k (score desc, id asc);
  2. all-gather of one (score, song, time) triple per query file per rank after each rank has run the
     sequence score over the candidates whose songs it owns; arg-max with the reference's tie rule
     (higher score, then lower song id, cpp/seqscore.cpp:121).

With world size 1 both collectives vanish and this is exactly ``Database.query_batch``.
The reference has no multi-GPU search (its faiss hook clones replicas, database.py:101-104).
"""
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
    """numpy statement of pfann_topk_merge: lists [G][Q][k] -> [Q][k], (valid, score desc, id asc)."""
    D = np.concatenate(list(dists), axis=1)
    I = np.concatenate(list(labels), axis=1)
    invalid = I < 0
    order = np.lexsort((I, -D.astype(np.float64), invalid), axis=1)[:, :k]
    return np.take_along_axis(D, order, 1), np.take_along_axis(I, order, 1)


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
    libpfann_b200 handle of this rank's shard (``pfann_b200.database.Database(..., songs=range)``)."""

    def __init__(self, backend, top_k, frame_shift_mul=1, hop_size=0.5, score_alpha=0.0, group=None):
        self.backend = backend
        self.top_k, self.fsm, self.hop_size, self.alpha = top_k, frame_shift_mul, hop_size, float(score_alpha)
        self.group = group
        import torch.distributed as dist
        self.dist = dist if (dist.is_available() and dist.is_initialized()) else None
        self.world = self.dist.get_world_size(group) if self.dist else 1

    def _all_gather(self, t):
        import torch
        if self.world == 1:
            return t.unsqueeze(0)
        out = torch.empty((self.world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        self.dist.all_gather(list(out.unbind(0)), t.contiguous(), group=self.group)   # nccl and gloo alike
        return out

    def query_batch(self, queries, query_index):
        """queries [sum len, d], query_index [nq, 2] (start, len), both replicated on every rank.
        Returns (score[nq], song[nq], time_s[nq]) identical on every rank."""
        d_l, i_l = self.backend.search_local(queries, self.top_k)             # tensors on the backend's device
        dg, ig = self._all_gather(d_l), self._all_gather(i_l)                 # exchange 1: [G, Q, k]
        D, I = self.backend.merge(dg, ig, self.top_k) if self.world > 1 else (d_l, i_l)
        s, g, t = self.backend.rerank_local(queries, query_index, I, self.top_k, self.fsm, self.alpha)
        import torch
        pack = torch.stack([s.double(), g.double(), t.double()], dim=1)       # exchange 2: [G, nq, 3]
        allp = self._all_gather(pack).cpu().numpy()
        score, song, tim = combine_best(allp[:, :, 0].astype(np.float32), allp[:, :, 1].astype(np.int64),
                                        allp[:, :, 2].astype(np.float32))
        return score, song, tim.astype(np.float64) * self.hop_size / self.fsm               # database.py:191


class GpuShard:
    """Per-rank engine over libpfann_b200: device-resident search / merge / rerank (all stream-ordered)."""

    def __init__(self, db):
        self.db = db          # pfann_b200.database.Database restricted to this rank's songs

    def search_local(self, queries, k):
        import torch
        from . import _lib
        dev = torch.device('cuda', self.db.device)
        q = torch.as_tensor(queries, dtype=torch.float32).to(dev).contiguous()
        D = torch.empty((q.shape[0], k), dtype=torch.float32, device=dev)
        I = torch.empty((q.shape[0], k), dtype=torch.int64, device=dev)
        _lib.use_torch_stream(self.db.device)
        _lib.check(_lib.lib().pfann_db_search(self.db.handle, _lib.ptr(q), q.shape[0], k, _lib.ptr(D), _lib.ptr(I)),
                   'pfann_db_search')
        self._q = q
        return D, I

    def merge(self, dg, ig, k):
        import torch
        from . import _lib
        G, Q = dg.shape[0], dg.shape[1]
        D = torch.empty((Q, k), dtype=torch.float32, device=dg.device)
        I = torch.empty((Q, k), dtype=torch.int64, device=dg.device)
        _lib.check(_lib.lib().pfann_topk_merge(_lib.ctx(self.db.device), _lib.ptr(dg.contiguous()),
                                               _lib.ptr(ig.contiguous()), G, Q, k, _lib.ptr(D), _lib.ptr(I)),
                   'pfann_topk_merge')
        return D, I

    def rerank_local(self, queries, query_index, labels, k, fsm, alpha):
        import torch
        from . import _lib
        dev = labels.device
        qi = torch.as_tensor(np.ascontiguousarray(query_index, dtype=np.int64)).to(dev)
        nq = qi.shape[0]
        s = torch.empty(nq, dtype=torch.float32, device=dev)
        g = torch.empty(nq, dtype=torch.int32, device=dev)
        t = torch.empty(nq, dtype=torch.float32, device=dev)
        _lib.check(_lib.lib().pfann_db_rerank(self.db.handle, _lib.ptr(self._q), _lib.ptr(qi), nq,
                                              _lib.ptr(labels.contiguous()), k, fsm, float(alpha), _lib.ptr(s),
                                              _lib.ptr(g), _lib.ptr(t)), 'pfann_db_rerank')
        return s, g, t
