"""CPU, world_size 2 over gloo: the host-side logic of the row-sharded search (pfann_b200/dist.py) -- shard
cuts at song boundaries, the all-gather of per-shard top-k, the merge and the winner combination -- with the
per-shard engine replaced by the oracle, against the unsharded oracle answer."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pfann_oracle as orc
from pfann_b200 import synth
from pfann_b200.dist import ShardedDatabase, combine_best, merge_topk, shard_songs


class OracleShard:
    """Stands in for GpuShard on a machine without a GPU: same contract, CPU oracle arithmetic."""

    def __init__(self, db, pos, songs):
        self.db, self.pos = db, pos
        self.s0, self.s1 = songs
        self.r0, self.r1 = int(pos[self.s0]), int(pos[self.s1])

    def search_local(self, queries, k):
        D, I = orc.flat_ip_search(self.db[self.r0:self.r1], np.asarray(queries), k)
        I = np.where(I >= 0, I + self.r0, -1)
        self._q = np.asarray(queries)
        return torch.from_numpy(D), torch.from_numpy(I)

    def merge(self, dg, ig, k):
        D, I = merge_topk(list(dg.numpy()), list(ig.numpy()), k)
        return torch.from_numpy(D), torch.from_numpy(I)

    def rerank_local(self, queries, query_index, labels, k, fsm, alpha):
        labels = labels.numpy()
        nq = len(query_index)
        s = np.full(nq, -np.inf, np.float32)
        g = np.full(nq, -1, np.int32)
        t = np.zeros(nq, np.float32)
        for i, (st, ln) in enumerate(query_index):
            lab = labels[st:st + ln].copy()
            lab[(lab < self.r0) | (lab >= self.r1)] = -1        # a shard scores only the songs it owns
            best, ss = orc.seq_score(self.db, self.pos, self._q[st:st + ln], lab, fsm, alpha)
            if best >= 0:
                # raw winner (no zero floor): highest per-song score, ties -> lower id
                cand = np.nonzero(ss[:, 0] > 0)[0]
                s[i], g[i], t[i] = (ss[best, 0], best, ss[best, 1]) if len(cand) else (0.0, best, 0.0)
        return torch.from_numpy(s), torch.from_numpy(g), torch.from_numpy(t)


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    db, key = synth.synth_db(4000, d=32, seed=3, song_len=37)
    pos = synth.song_pos_from_key(key)
    qs, songs, offs = synth.synth_queries(db, key, 6, q_len=19, seed=5)
    q = qs.reshape(-1, 32)
    qi = np.stack([np.arange(6) * 19, np.full(6, 19)], 1).astype(np.int64)
    shard = OracleShard(db, pos, shard_songs(pos, world)[rank])
    sdb = ShardedDatabase(shard, 10, 1, 0.5)
    score, song, tim = sdb.query_batch(q, qi)
    if rank == 0:
        ret['score'], ret['song'], ret['time'] = score, song, tim
    dist.destroy_process_group()


def test_sharded_search_world2_matches_unsharded():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    db, key = synth.synth_db(4000, d=32, seed=3, song_len=37)
    pos = synth.song_pos_from_key(key)
    qs, songs, offs = synth.synth_queries(db, key, 6, q_len=19, seed=5)
    for i in range(6):
        _, labels = orc.flat_ip_search(db, qs[i], 10)
        sco, (sid, tim), _ = orc.query_embeddings_cpp(db, pos, qs[i], labels, 1, 0.5)
        assert ret['song'][i] == sid == songs[i]
        assert ret['time'][i] == tim == offs[i] * 0.5
        assert ret['score'][i] == np.float32(sco)


def test_shard_songs_cuts_at_song_boundaries():
    key = np.array([30, 0, 5, 59, 19, 40, 1, 25, 59, 33, 12, 59], np.int32)
    pos = synth.song_pos_from_key(key)
    for world in (1, 2, 3, 4, 8, 16):
        sh = shard_songs(pos, world)
        assert sh[0][0] == 0 and sh[-1][1] == len(key)
        assert all(a[1] == b[0] for a, b in zip(sh, sh[1:]))
        assert all(a[0] <= a[1] for a in sh)


def test_merge_topk_and_combine_best_rules():
    d1 = np.array([[0.9, 0.5, -3.4e38]], np.float32)
    i1 = np.array([[7, 3, -1]], np.int64)
    d2 = np.array([[0.9, 0.7, 0.1]], np.float32)
    i2 = np.array([[5, 100, 101]], np.int64)
    D, I = merge_topk([d1, d2], [i1, i2], 4)
    assert list(I[0]) == [5, 7, 100, 3]                      # equal scores -> lower id first, -1 never wins
    s, g, t = combine_best([[0.5, -1.0, 0.2]], [[3, 2, -1]], [[4.0, 1.0, 0.0]])
    assert list(g) == [3, 2, -1] and list(s) == [0.5, 0.0, 0.0] and list(t) == [4.0, 0.0, 0.0]
    s, g, t = combine_best([[0.5], [0.5]], [[9], [4]], [[1.0], [2.0]])
    assert g[0] == 4 and t[0] == 2.0                          # tie -> lower song id (seqscore.cpp:121)
