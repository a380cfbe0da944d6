"""GPU (-m gpu): BASELINE-size checks through size-independent properties (the oracle would take minutes here):
config 3 shape (1M x 128 database, 19-vector query files, top-20) and a config-2 slice of the extractor."""
import numpy as np
import pytest

torch = pytest.importorskip('torch')
from pfann_b200 import synth  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def big_db():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from pfann_b200.database import Database
    g = torch.Generator(device='cuda')
    g.manual_seed(123)
    n, d, song_len = 1_000_000, 128, 59
    emb = torch.randn((n, d), generator=g, device='cuda')
    emb = emb / emb.norm(dim=1, keepdim=True)
    n_songs = (n + song_len - 1) // song_len
    key = np.full(n_songs, song_len, np.int32)
    key[-1] = n - song_len * (n_songs - 1)
    db = Database.from_arrays(emb, key, {'top_k': 20, 'frame_shift_mul': 1}, 0.5, device=0)
    return db, emb, key


def test_search_properties_1m(big_db):
    db, emb, key = big_db
    rows = torch.tensor([0, 1, 58, 59, 123456, 999_998, 999_999], device='cuda')
    q = emb[rows].cpu().numpy()
    D, I = db.search(q)
    assert (I[:, 0] == rows.cpu().numpy()).all()                       # a row is its own nearest neighbour
    np.testing.assert_allclose(D[:, 0], 1.0, atol=2e-6)
    assert (np.diff(D, axis=1) <= 0).all()                             # sorted descending
    assert all(len(set(r)) == 20 for r in I) and (I >= 0).all() and (I < 1_000_000).all()
    # exactness: every returned distance is the canonical fp32 inner product of that row (k-sequential fma)
    e = emb[torch.from_numpy(I.reshape(-1)).cuda()].cpu().numpy().reshape(7, 20, 128)
    acc = np.zeros((7, 20), np.float32)
    for kk in range(128):
        acc = np.float32(np.float64(e[:, :, kk]) * np.float64(q[:, None, kk]) + np.float64(acc))   # fma = 1 rounding
    assert np.array_equal(acc.view(np.uint32), D.view(np.uint32))
    # linearity: scaling the query by 2 scales the distances by exactly 2 and keeps the labels
    D2, I2 = db.search(2.0 * q)
    assert np.array_equal(I2, I) and np.array_equal(D2, 2.0 * D)
    # completeness: nothing outside the returned set beats the k-th score (checked on one query by brute force)
    s = (emb @ torch.from_numpy(q[4]).cuda()).cpu().numpy()
    assert (np.sort(s)[-20:][::-1] - D[4]).max() < 1e-6
    assert set(np.argsort(-s, kind='stable')[:20]) == set(I[4])


def test_planted_queries_1m(big_db):
    """config 3: 2k-style noisy query files (here 200) -> the planted (song, offset) comes back for every one,
    batched and one by one give identical answers."""
    db, emb, key = big_db
    rng = np.random.Generator(np.random.PCG64(7))
    nq, q_len = 200, 19
    songs = rng.integers(0, len(key) - 1, nq)
    offs = rng.integers(0, 59 - q_len + 1, nq)
    idx = (songs * 59 + offs)[:, None] + np.arange(q_len)[None, :]
    q = emb[torch.from_numpy(idx.reshape(-1)).cuda()].cpu().numpy().reshape(nq, q_len, 128)
    q = q + rng.standard_normal(q.shape, dtype=np.float32) * np.float32(1.0 / np.sqrt(128))
    q /= np.linalg.norm(q, axis=2, keepdims=True)
    qi = np.stack([np.arange(nq) * q_len, np.full(nq, q_len)], 1).astype(np.int64)
    score, song, tim, _ = db.query_batch(q.reshape(-1, 128), qi)
    assert (song == songs).all() and (tim == offs * 0.5).all()
    assert (score > 0.3).all() and (score <= 1.0).all()
    for i in (0, 57, 199):
        sco, (sid, t), ss = db.query_embeddings(q[i])
        assert sid == song[i] and t == tim[i] and np.float32(sco) == score[i]
        assert ss[sid, 0] == np.float32(sco) and (ss[:, 0] <= np.float32(sco)).all()


def test_extract_slice_properties():
    """config 2 slice: 40 clips x 30 s -> 2360 fingerprints; unit norm, run-to-run identical, and independent of
    how the clips are batched into calls."""
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from pfann_b200.extract import Extractor
    params = synth.read_config('default')
    ex = Extractor(params, synth.make_state_dict(params, seed=11), device=0, precision='bf16', chunk=512)
    clip = 240000
    pcm = np.concatenate([synth.synth_pcm(i, clip) for i in range(40)])
    off = np.arange(41, dtype=np.int64) * clip
    z, counts = ex.extract_pcm16(pcm, off)
    assert z.shape == (2360, 128) and (counts == 59).all()
    np.testing.assert_allclose(np.linalg.norm(z, axis=1), 1.0, atol=1e-5)
    z2, _ = ex.extract_pcm16(pcm, off)
    assert np.array_equal(z, z2)
    za, _ = ex.extract_pcm16(pcm[:clip * 7], off[:8])
    zb, _ = ex.extract_pcm16(pcm[clip * 7:], off[7:] - off[7])
    assert np.array_equal(np.concatenate([za, zb]), z)
    # neighbouring segments of the same clip overlap by half a second: closer than segments of other clips
    same = (z[0:58] * z[1:59]).sum(1).mean()
    other = (z[0:58] * z[59:117]).sum(1).mean()
    assert same > other


def test_extract_large_chunk_equals_small_chunk():
    """The bench configuration (chunk = 16384 segments: hundreds of tiles per CTA position in the fused conv+LayerNorm
    kernel, statistics exchanged between up to 16 CTAs per sample) gives bit for bit the fingerprints of a small-chunk
    run: per-sample statistics are summed in a fixed order, independent of how segments are batched."""
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from pfann_b200.extract import Extractor
    params = synth.read_config('default')
    sd = synth.make_state_dict(params, seed=11)
    big = Extractor(params, sd, device=0, precision='bf16', chunk=16384)
    small = Extractor(params, sd, device=0, precision='bf16', chunk=96)
    clip, n_clips = 240000, 290                                   # 17110 segments: one full chunk + a ragged one
    g = torch.Generator(device='cuda')
    g.manual_seed(5)
    pcm = (torch.randn(n_clips * clip, generator=g, device='cuda') * 4000).to(torch.int16)
    off = np.arange(n_clips + 1, dtype=np.int64) * clip
    z, counts = big.extract_pcm16(pcm, off)
    assert z.shape == (n_clips * 59, 128) and (counts == 59).all()
    torch.testing.assert_close(z.norm(dim=1), torch.ones(z.shape[0], device='cuda'), atol=1e-5, rtol=0)
    for c0 in (0, 137, 277, 286):                                 # first chunk, middle, chunk boundary, ragged tail
        zs, _ = small.extract_pcm16(pcm[c0 * clip:(c0 + 4) * clip], off[c0:c0 + 5] - off[c0])
        assert torch.equal(zs, z[c0 * 59:(c0 + 4) * 59]), c0
