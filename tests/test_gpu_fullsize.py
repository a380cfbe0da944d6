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


def test_end_to_end_noisy_pcm_queries_bf16_equals_fp32_reference_path():
    """north_star: "song-id / segment-offset pairs are bit-exact on the same inputs".  From PCM to the answer:
    1 700 clips x 30 s -> a 100 300-row database; 240 query files = 10 s excerpts with additive noise (10 dB SNR),
    re-quantised to int16.  Product path: bf16 tensor-core fingerprints -> GPU search + sequence score.  Reference
    path: fp32 fingerprints (the validation-grade CUDA-core path, itself tied to the double-precision oracle here
    on two query files) -> exact CPU inner-product top-k -> the oracle's seq_score (restatement of cpp/seqscore.cpp).
    Both must name the same (song, offset) for every query, and that must be the excerpt's true position."""
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from oracle import pfann_oracle as orc          # checker only
    from pfann_b200.database import Database
    from pfann_b200.extract import Extractor
    params = synth.read_config('default')
    sd = synth.make_state_dict(params, seed=11)
    n_clips, clip, hop, q_samples, k = 1700, 240000, 4000, 80000, 20
    g = torch.Generator(device='cuda')
    g.manual_seed(2026)
    t = torch.arange(clip, device='cuda', dtype=torch.float32) / 8000.0
    pcm = torch.empty(n_clips * clip, dtype=torch.int16, device='cuda')
    for c0 in range(0, n_clips, 100):
        w = torch.randn((100, clip + 1), generator=g, device='cuda')
        x = 0.6 * w[:, 1:] + 0.4 * w[:, :-1]
        for _ in range(3):
            f = 300.0 + 3600.0 * torch.rand((100, 1), generator=g, device='cuda')
            a = 0.3 + 0.7 * torch.rand((100, 1), generator=g, device='cuda')
            ph = 6.2831853 * torch.rand((100, 1), generator=g, device='cuda')
            x = x + a * torch.sin(6.2831853 * f * t[None, :] + ph)
        x = x * (0.5 / x.abs().amax(dim=1, keepdim=True))
        pcm[c0 * clip:(c0 + 100) * clip] = torch.round(x * 32767.0).to(torch.int16).reshape(-1)
    off = np.arange(n_clips + 1, dtype=np.int64) * clip
    ex16 = Extractor(params, sd, device=0, precision='bf16', chunk=8192)
    ex32 = Extractor(params, sd, device=0, precision='fp32', chunk=512)
    db16, counts = ex16.extract_pcm16(pcm, off)
    db32, _ = ex32.extract_pcm16(pcm, off)
    assert db16.shape[0] == n_clips * 59 >= 100_000 and (counts == 59).all()
    # queries: noisy excerpts
    rng = np.random.Generator(np.random.PCG64(99))
    nq = 240
    songs = rng.integers(0, n_clips, nq)
    offs = rng.integers(0, (clip - q_samples) // hop + 1, nq)
    qpcm = torch.empty(nq * q_samples, dtype=torch.int16, device='cuda')
    for i in range(nq):
        s0 = int(songs[i]) * clip + int(offs[i]) * hop
        x = pcm[s0:s0 + q_samples].float()
        noise = torch.randn(q_samples, generator=g, device='cuda') * (x.std() * 10 ** (-10 / 20))
        qpcm[i * q_samples:(i + 1) * q_samples] = torch.clamp(torch.round(0.7 * (x + noise)), -32768, 32767).to(torch.int16)
    qoff = np.arange(nq + 1, dtype=np.int64) * q_samples
    q16, qc = ex16.extract_pcm16(qpcm, qoff)
    q32, _ = ex32.extract_pcm16(qpcm, qoff)
    assert (qc == 19).all()
    cos = (q16 * q32).sum(1)
    assert float((1 - cos).max()) < 1e-3                                         # bf16 vs fp32 fingerprints
    # the fp32 path is anchored to the double-precision oracle on two query files (38 segments)
    rows = np.concatenate([orc.frame_pcm16(qpcm[i * q_samples:(i + 1) * q_samples].cpu().numpy(), 8000, hop) for i in (0, 1)])
    zo = orc.fpnetwork_forward(sd, orc.melspec(rows, params), params)
    assert (1 - (zo * q32[:38].cpu().numpy()).sum(1)).max() < 2e-5
    # product path
    key = counts.astype(np.int32)
    dbo = Database.from_arrays(db16, key, {'top_k': k, 'frame_shift_mul': 1}, 0.5, device=0)
    qi = np.stack([np.arange(nq) * 19, np.full(nq, 19)], 1).astype(np.int64)
    score, song, tim, _ = dbo.query_batch(q16.cpu().numpy(), qi)
    # reference path on the CPU
    pos = synth.song_pos_from_key(key)
    dbn, qn = db32.cpu().numpy(), q32.cpu().numpy()
    ref_song, ref_time = np.empty(nq, np.int64), np.empty(nq, np.float64)
    for i in range(nq):
        qq = qn[i * 19:(i + 1) * 19]
        sc = qq @ dbn.T
        lab = np.argsort(-sc, axis=1, kind='stable')[:, :k].astype(np.int64)
        best, ss = orc.seq_score(dbn, pos, qq, lab, 1, 0.0)
        ref_song[i], ref_time[i] = best, ss[best, 1] * 0.5
    assert np.array_equal(song, ref_song) and np.array_equal(tim, ref_time)
    assert np.array_equal(ref_song, songs) and np.array_equal(ref_time, offs * 0.5)
    assert float(score.min()) > 0.2 and float(score.max()) < 0.999                 # the noise is felt
