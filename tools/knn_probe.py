"""GPU probe: time the kNN scan kernel alone on a big database under the PFANN_KNN_DEBUG knobs
(0 full, 1 no filter, 2 no TMEM loads, 3 no MMAs = pure TMA stream) to see which stage paces it."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pfann_b200 import _lib  # noqa: E402
from pfann_b200.database import Database  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dev = torch.device('cuda', 0)
emb = torch.randn((n, 128), device=dev)
emb /= emb.norm(dim=1, keepdim=True)
key = np.full(n // 50, 50, np.int32)
db = Database.from_arrays(emb, key, {'top_k': 20}, 0.5, device=0)
del emb
prof = torch.zeros((148, 8), dtype=torch.int64, device=dev)
os.environ['PFANN_KNN_PROF_PTR'] = str(prof.data_ptr())
for Q in (19, 256):
    q = torch.randn((Q, 128), device=dev)
    q /= q.norm(dim=1, keepdim=True)
    D = torch.empty((Q, 20), device=dev)
    I = torch.empty((Q, 20), dtype=torch.int64, device=dev)
    for dbg in (0, 1, 2, 3):
        os.environ['PFANN_KNN_DEBUG'] = str(dbg)
        prof.zero_()
        _lib.use_torch_stream(0)
        for it in range(3):
            if it == 1:
                _lib.profile(0, True)
            _lib.check(_lib.lib().pfann_db_search(db.handle, _lib.ptr(q), Q, 20, _lib.ptr(D), _lib.ptr(I)))
        p = _lib.profile_read(0)
        _lib.profile(0, False)
        ms, cnt = p['knn_scan']
        sel = p['knn_select'][0]
        print('Q=%3d debug=%d  scan %.3f ms over %d launches (%.1f GB/s of bf16 DB per full pass)  select %.3f ms'
              % (Q, dbg, ms / 2, cnt // 2, n * 256 / (ms / 2 / 1e3) / 1e9 if ms else 0, sel / 2), flush=True)
        pr = prof.cpu().numpy().astype(np.float64)
        tiles = n / 128 / 148 * 3  # 3 searches x (prepass + full scan) accumulate; per-CTA tiles of the full scans dominate
        print('     cycles/tile/CTA: producer wait empty %.0f | mma wait tempty %.0f, wait full %.0f | epilogue (avg of 4 warps) wait tfull %.0f, work %.0f'
              % tuple(pr[:, j].mean() / tiles / (4 if j >= 3 else 1) for j in range(5)), flush=True)
