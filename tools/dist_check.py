"""torchrun --nproc-per-node N tools/dist_check.py : the row-sharded search (NCCL all-gather of per-shard top-k,
per-shard rerank, winner exchange) gives on every rank exactly the answers of the unsharded single-GPU database."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pfann_b200 import synth  # noqa: E402
from pfann_b200.database import Database  # noqa: E402
from pfann_b200.dist import GpuShard, ShardedDatabase, shard_songs  # noqa: E402

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
ok = True
for (n, k, fsm) in ((200_000, 20, 1), (50_000, 100, 2)):
    db, key = synth.synth_db(n, d=128, seed=4, song_len=59)
    pos = synth.song_pos_from_key(key)
    qs, songs, offs = synth.synth_queries(db, key, 64, q_len=19 * fsm, seed=8)
    if fsm > 1:   # finer query hop: repeat every database step fsm times along the diagonal is not needed for parity
        pass
    q = qs.reshape(-1, 128)
    qi = np.stack([np.arange(64) * qs.shape[1], np.full(64, qs.shape[1])], 1).astype(np.int64)
    shard = Database.from_arrays(db, key, {'top_k': k, 'frame_shift_mul': fsm}, 0.5, device=local,
                                 songs=shard_songs(pos, world)[rank])
    sdb = ShardedDatabase(GpuShard(shard), k, fsm, 0.5)
    score, song, tim = sdb.query_batch(q, qi)
    s2, g2, t2 = sdb.query_batches(q, qi, 24)          # three batches, one read-back
    full = Database.from_arrays(db, key, {'top_k': k, 'frame_shift_mul': fsm}, 0.5, device=local)
    rs, rg, rt, _ = full.query_batch(q, qi)
    same = np.array_equal(song, rg) and np.array_equal(tim, rt) and np.array_equal(score, rs)
    same = same and np.array_equal(song, g2) and np.array_equal(tim, t2) and np.array_equal(score, s2)
    if fsm == 1:
        same = same and np.array_equal(song, songs) and np.array_equal(tim, offs * 0.5)
    # the merged top-k equals the unsharded top-k bit for bit
    D, I = full.search(q[:57], k)
    b = sdb.backend
    qd = b.to_device(q[:57])
    thr = b.thresholds(qd, k)
    if world > 1:
        dist.all_reduce(thr, op=dist.ReduceOp.MAX)
    Dm, Im = b.merge_keys(sdb._all_gather(b.filtered_keys(qd, k, thr, False)), k, want_dist=True)
    same = same and np.array_equal(Im.cpu().numpy(), I) and np.array_equal(Dm.cpu().numpy().view(np.uint32), D.view(np.uint32))
    t = torch.tensor([int(same)], device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    ok = ok and bool(t.item())
    if rank == 0:
        print('n=%d k=%d fsm=%d world=%d: sharded == unsharded on all ranks: %s' % (n, k, fsm, world, bool(t.item())), flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
