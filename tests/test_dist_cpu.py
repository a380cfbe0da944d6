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
from pfann_b200.dist import ShardedDatabase, combine_best, merge_topk, pack_keys, shard_songs, unpack_keys


class OracleShard:
    """Stands in for GpuShard on a machine without a GPU: same contract, CPU oracle arithmetic."""

    def __init__(self, db, pos, songs):
        self.db, self.pos = db, pos
        self.s0, self.s1 = songs
        self.r0, self.r1 = int(pos[self.s0]), int(pos[self.s1])
        self._norm = float(np.linalg.norm(db[self.r0:self.r1], axis=1).max()) if self.r1 > self.r0 else 0.0

    def torch_device(self):
        return torch.device('cpu')

    def to_device(self, queries):
        return torch.as_tensor(np.asarray(queries), dtype=torch.float32)

    def max_norm(self):
        return self._norm

    def set_max_norm(self, v):
        assert v >= self._norm
        self._norm = v

    def thresholds(self, q, k):
        # like the GPU pre-pass: the k-th best score of a SAMPLE of the shard (a lower bound of its k-th best)
        rows = self.db[self.r0:self.r1][::3]
        sc = np.sort(q.numpy() @ rows.T, axis=1)[:, ::-1]
        thr = sc[:, k - 1] if rows.shape[0] >= k else np.full(q.shape[0], -np.inf, np.float32)
        return torch.from_numpy(np.ascontiguousarray(thr, dtype=np.float32))

    def sample_topk(self, q, k):
        # the k best SAMPLED scores per query as opaque sortable keys (here: the fp32 bits of non-negative shifted scores)
        rows = self.db[self.r0:self.r1][::3]
        sc = np.sort(q.numpy() @ rows.T, axis=1)[:, ::-1][:, :k].astype(np.float32)
        out = np.full((q.shape[0], k), -np.inf, np.float32)
        out[:, :sc.shape[1]] = sc
        return torch.from_numpy(out.view(np.int32).copy())

    def thresholds_from_topk(self, top_g, q, k):
        sc = np.concatenate(list(top_g.numpy().view(np.float32)), axis=1)      # [Q, G * k]
        return torch.from_numpy(np.ascontiguousarray(np.sort(sc, axis=1)[:, ::-1][:, k - 1]))

    def take_overflow(self):
        return 0

    def filtered_keys(self, q, k, thr, defer=False):
        D, I = orc.flat_ip_search(self.db[self.r0:self.r1], q.numpy(), k)
        keep = (I >= 0) & (D >= thr.numpy()[:, None])              # rows below the (global) threshold are dropped
        I = np.where(keep, I + self.r0, -1)
        return torch.from_numpy(pack_keys(D, I).view(np.int64))

    def merge_keys(self, keys_g, k):
        allk = np.concatenate(list(keys_g.numpy().view(np.uint64)), axis=1)
        srt = np.sort(allk, axis=1)[:, ::-1][:, :k]                 # descending unsigned keys = score desc, id asc
        return torch.from_numpy(np.ascontiguousarray(unpack_keys(srt)[1]))

    def rerank_packed(self, q, query_index, labels, k, fsm, alpha):
        labels, qn = labels.numpy(), q.numpy()
        nq = len(query_index)
        out = np.zeros((nq, 4), np.float32)
        song = np.full(nq, -1, np.int32)
        out[:, 0] = -np.inf
        for i, (st, ln) in enumerate(query_index):
            lab = labels[st:st + ln].copy()
            lab[(lab < self.r0) | (lab >= self.r1)] = -1        # a shard scores only the songs it owns
            best, ss = orc.seq_score(self.db, self.pos, qn[st:st + ln], lab, fsm, alpha)
            if best >= 0:
                # raw winner (no zero floor): highest per-song score, ties -> lower id
                cand = np.nonzero(ss[:, 0] > 0)[0]
                out[i, 0], song[i], out[i, 2] = (ss[best, 0], best, ss[best, 1]) if len(cand) else (0.0, best, 0.0)
        out[:, 1] = song.view(np.float32)
        return torch.from_numpy(out)

    def combine(self, packed_g):
        p = packed_g.numpy()
        s, g, t = combine_best(p[:, :, 0], np.ascontiguousarray(p[:, :, 1]).view(np.int32), p[:, :, 2])
        out = np.zeros((p.shape[1], 4), np.float32)
        out[:, 0], out[:, 1], out[:, 2] = s, g.astype(np.int32).view(np.float32), t
        return torch.from_numpy(out)


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
    # the union-of-samples threshold is a lower bound of the global k-th best score and at least as tight as the
    # maximum of the per-shard thresholds (the older exchange, still used by backends without sample_topk)
    qt = shard.to_device(q)
    thr_union = shard.thresholds_from_topk(sdb._all_gather(shard.sample_topk(qt, 10)), qt, 10)
    thr_max = shard.thresholds(qt, 10)
    dist.all_reduce(thr_max, op=dist.ReduceOp.MAX)
    kth = np.sort(q @ db.T, axis=1)[:, ::-1][:, 9]
    assert (thr_union.numpy() <= kth + 1e-6).all() and (thr_union.numpy() >= thr_max.numpy() - 1e-6).all()
    score, song, tim = sdb.query_batch(q, qi)
    s2, g2, t2 = sdb.query_batches(q, qi, 4)                      # two batches, one read-back: same answers
    assert np.array_equal(score, s2) and np.array_equal(song, g2) and np.array_equal(tim, t2)
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


def test_key_packing_orders_like_score_desc_id_asc():
    D = np.array([[0.5, -0.25, 0.5, 0.0, -0.0, 1e-30]], np.float32)
    I = np.array([[9, 4, 3, 7, 8, -1]], np.int64)
    keys = pack_keys(D, I)
    order = np.argsort(keys[0], kind='stable')[::-1]                 # unsigned integer order, descending
    assert list(I[0][order]) == [3, 9, 7, 8, 4, -1]                  # +0.0 sorts above -0.0 (distinct bit patterns)
    srt = np.sort(keys, axis=1)[:, ::-1]
    d2, i2 = unpack_keys(srt)
    assert list(i2[0][:2]) == [3, 9] and i2[0][-1] == -1 and d2[0][0] == np.float32(0.5)
    assert np.array_equal(unpack_keys(keys)[0][0][:5].view(np.uint32), D[0][:5].view(np.uint32))   # bit-exact scores


def _grad_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import torch
    from pfann_b200.train import allreduce_gradients
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 3))
    x = torch.arange(8 * 5, dtype=torch.float32).reshape(8, 5) / 40
    net(x[rank * 4:(rank + 1) * 4]).pow(2).sum().backward()          # each rank: its rows of a sum-type loss
    net[2].bias.grad = None                                          # parameters without a gradient are skipped
    allreduce_gradients(list(net.parameters()))
    if rank == 0:
        ret['grads'] = [None if p.grad is None else p.grad.numpy().copy() for p in net.parameters()]
    dist.destroy_process_group()


def test_gradient_sum_over_ranks_equals_the_whole_batch():
    """train step at world 2 (BASELINE configs[4]): one flat all-reduce(SUM) of the per-rank gradients of a loss that
    is a sum over rows gives the single-process gradient on all rows."""
    import torch
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_grad_worker, args=(2, port, ret), nprocs=2, join=True)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 3))
    x = torch.arange(8 * 5, dtype=torch.float32).reshape(8, 5) / 40
    net(x).pow(2).sum().backward()
    got = ret['grads']
    assert got[3] is None
    for g, p in zip(got[:3], list(net.parameters())[:3]):
        np.testing.assert_allclose(g, p.grad.numpy(), rtol=1e-5, atol=1e-6)
