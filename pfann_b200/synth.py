"""Seeded synthetic inputs for the hot path: model weights, 8 kHz PCM, fingerprint databases.

There are no datasets or checkpoints offline, so tests and ``bench.py`` build everything from
seeds.  Generators use numpy's PCG64 stream (stable across numpy versions) so golden fixtures
made in the build container reproduce on the GPU box.  Shapes follow SURVEY.md section 8(d).
"""
import json
import os

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def read_config(name_or_path):
    """JSON hyper-parameters, same reader semantics as the reference (simpleutils.py:30-32)."""
    path = name_or_path
    if not os.path.exists(path):
        path = os.path.join(REPO, 'configs', name_or_path if name_or_path.endswith('.json')
                            else name_or_path + '.json')
    with open(path, 'r') as fin:
        return json.load(fin)


def model_dims(params):
    """(d, h, u, F, T) exactly as builder.py:46-51 derives them."""
    m = params['model']
    segn = int(params['segment_size'] * params['sample_rate'])
    T = (segn + params['stft_hop'] - 1) // params['stft_hop']
    return m['d'], m['h'], m['u'], params['n_mels'], T


def layer_shapes(params):
    """(Cin, Cout, F, T) per SeparableConv2d (model.py:79-93)."""
    d, h, u, F, T = model_dims(params)
    ch = [1, d, d, 2 * d, 2 * d, 4 * d, 4 * d, h, h]
    out = []
    for i, (st, sf) in enumerate(layer_strides(params)):
        out.append((ch[i], ch[i + 1], F, T))
        F = (F - 1) // sf + 1
        T = (T - 1) // st + 1
    if F != 1 or T != 1:
        raise ValueError('output must be 1x1')  # model.py:94
    return out


def layer_strides(params):
    """(time stride of conv1, frequency stride of conv2) per SeparableConv2d (model.py:82-85)."""
    st = params['model'].get('strides')
    if st is None:
        return [(2, 2)] * 8
    return [(int(st[i][0][1]), int(st[i][1][0])) for i in range(8)]


def make_state_dict(params, seed=0):
    """Random FpNetwork weights keyed like the reference's ``model.pt`` (model.py).

    conv/linear weights ~ N(0, 1/fan_in); LayerNorm affine is deliberately non-trivial
    (gamma = 1 + 0.1 N, beta = 0.1 N) so that the per-element (C,F,T) indexing is exercised.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    m = params['model']
    d, h, u, F, T = model_dims(params)
    fuller = bool(m.get('fuller', False))
    sd = {}

    def nrm(shape, std):
        return (rng.standard_normal(shape, dtype=np.float32) * np.float32(std)).astype(np.float32)

    strides = layer_strides(params)
    for l, (ci, co, f, t) in enumerate(layer_shapes(params)):
        f2, t2 = (f - 1) // strides[l][1] + 1, (t - 1) // strides[l][0] + 1
        p = 'f.convs.%d.' % l
        sd[p + 'conv1.weight'] = nrm((co, ci, 1, 3), (1.0 / (3 * ci)) ** 0.5)
        sd[p + 'conv1.bias'] = nrm((co,), 0.1)
        sd[p + 'ln1.weight'] = (1 + nrm((co, f, t2), 0.1)).astype(np.float32)
        sd[p + 'ln1.bias'] = nrm((co, f, t2), 0.1)
        cin2 = co if fuller else 1
        sd[p + 'conv2.weight'] = nrm((co, cin2, 3, 1), (1.0 / (3 * cin2)) ** 0.5)
        sd[p + 'conv2.bias'] = nrm((co,), 0.1)
        sd[p + 'ln2.weight'] = (1 + nrm((co, f2, t2), 0.1)).astype(np.float32)
        sd[p + 'ln2.bias'] = nrm((co, f2, t2), 0.1)
    v = h // d
    sd['g.linear1.weight'] = nrm((d * u, v, 1), (1.0 / v) ** 0.5)
    sd['g.linear1.bias'] = nrm((d * u,), 0.1)
    sd['g.linear2.weight'] = nrm((d, u, 1), (1.0 / u) ** 0.5)
    sd['g.linear2.bias'] = nrm((d,), 0.1)
    return sd


def synth_pcm(clip_id, n_samples, sample_rate=8000):
    """int16 mono clip: low-passed ("pink-ish") noise + 3 sinusoids in 300..4000 Hz, peak 0.5 FS."""
    rng = np.random.Generator(np.random.PCG64(1000 + int(clip_id)))
    w = rng.standard_normal(n_samples + 1).astype(np.float32)
    x = 0.6 * w[1:] + 0.4 * w[:-1]
    t = np.arange(n_samples, dtype=np.float64) / sample_rate
    for _ in range(3):
        f = rng.uniform(300.0, 3900.0)
        a = rng.uniform(0.3, 1.0)
        ph = rng.uniform(0, 2 * np.pi)
        x = x + (a * np.sin(2 * np.pi * f * t + ph)).astype(np.float32)
    x = x * (0.5 / np.max(np.abs(x)))
    return np.round(x * 32767.0).astype(np.int16)


def synth_segments(n, seed=0, seg=8000):
    """[n, seg] fp32 zero-mean rows the way MusicDataset hands them over (musicdata.py:87-88)."""
    out = np.empty((n, seg), np.float32)
    for i in range(n):
        x = synth_pcm(seed * 7919 + i, seg).astype(np.float32) * np.float32(1 / 32768)
        out[i] = x - x.mean(dtype=np.float32)
    return out


def synth_db(n, d=128, seed=0, song_len=59):
    """Unit-norm fp32 fingerprints grouped into songs of ``song_len`` rows (last one shorter).

    Returns (db[n,d], landmark_key int32[n_songs]) -- the ``embeddings`` / ``landmarkKey`` files of
    builder.py:99,138-139."""
    rng = np.random.Generator(np.random.PCG64(seed))
    db = rng.standard_normal((n, d), dtype=np.float32)
    db /= np.linalg.norm(db, axis=1, keepdims=True)
    n_songs = (n + song_len - 1) // song_len
    key = np.full(n_songs, song_len, np.int32)
    key[-1] = n - song_len * (n_songs - 1)
    return db, key


def song_pos_from_key(key):
    """database.py:83-86."""
    return np.pad(np.cumsum(np.asarray(key), dtype=np.int64), (1, 0))


def synth_queries(db, key, n_queries, q_len=19, noise=1.0, seed=7):
    """Each query = a real database diagonal (random song, random offset) + Gaussian noise whose
    per-vector norm is ~``noise`` (per-element std noise/sqrt(d)), renormalised.

    Returns (queries[n_queries, q_len, d], truth_song[n_queries], truth_offset[n_queries])."""
    rng = np.random.Generator(np.random.PCG64(seed))
    pos = song_pos_from_key(key)
    d = db.shape[1]
    ok = np.nonzero(np.asarray(key) >= q_len)[0]
    songs = ok[rng.integers(0, len(ok), n_queries)]
    offs = np.array([rng.integers(0, key[s] - q_len + 1) for s in songs], np.int64)
    idx = (pos[songs] + offs)[:, None] + np.arange(q_len)[None, :]
    q = db[idx] + rng.standard_normal((n_queries, q_len, d), dtype=np.float32) * np.float32(noise / np.sqrt(d))
    q /= np.linalg.norm(q, axis=2, keepdims=True)
    return q.astype(np.float32), songs.astype(np.int64), offs


# NAF-style stride schedule for F = 256, T = 32 (tools/gen_golden.py MODEL_VARIANTS['strides'])
NAF_STRIDES = [[[1, 2], [2, 1]], [[1, 2], [2, 1]], [[1, 2], [2, 1]], [[1, 2], [2, 1]],
               [[1, 1], [2, 1]], [[1, 2], [2, 1]], [[1, 1], [2, 1]], [[1, 1], [2, 1]]]
# training-step cases: name -> (config, model option overrides, batch, weight seed)
TRAIN_CASES = {
    'tiny': ('tiny', {}, 8, 31),
    'tiny_dw': ('tiny', {'fuller': False}, 8, 32),
    'tiny_strides': ('tiny', {'strides': NAF_STRIDES}, 6, 33),
    'n640d64': ('n640d64', {}, 4, 34),
    'tiny_elu': ('tiny', {'conv_activation': 'ELU'}, 6, 35),
    'tiny_act_first': ('tiny', {'relu_after_bn': False}, 6, 36),
    'tiny_elu_act_first': ('tiny', {'conv_activation': 'ELU', 'relu_after_bn': False, 'fuller': False}, 6, 37),
}


def train_case_input(name):
    """Seeded stand-in for a batch of log-mel segments [B][256][32] (pairs 2i, 2i+1 correlated like a clip and its
    augmentation); shared by tools/gen_golden.py and the tests."""
    cfg, opt, B, seed = TRAIN_CASES[name]
    rng = np.random.Generator(np.random.PCG64(1000 + seed))
    x = rng.standard_normal((B, 256, 32)).astype(np.float32) * np.float32(0.5)
    x[1::2] = x[0::2] + np.float32(0.3) * x[1::2]
    return x


def grad_sample_index(numel, seed):
    """Which elements of a gradient the fixture keeps (all of a small one, 512 seeded positions of a large one)."""
    if numel <= 512:
        return np.arange(numel)
    return np.sort(np.random.Generator(np.random.PCG64(seed)).choice(numel, 512, replace=False))
