"""ctypes front end of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this module; the product package ``pfann_b200`` never does.

Citations are relative to /root/reference.  See ``pfann_oracle.c`` for the C restatements;
this file adds the numpy restatement of the Python rerank (``database.py:117-166``) and the
binding to the reference's own ``cpp/seqscore.cpp`` compiled into ``oracle/_ref/``.
"""
import ctypes
import os
import subprocess
from ctypes import POINTER, c_double, c_float, c_int, c_int16, c_int64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None


def build(force=False):
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    so = os.path.join(_HERE, 'libpfann_oracle.so')
    src = os.path.join(_HERE, 'pfann_oracle.c')
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', _HERE, '-s', 'all'], stdout=subprocess.DEVNULL)
    return so


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t):
    return a.ctypes.data_as(POINTER(t))


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        L.orc_version.restype = ctypes.c_longlong
        L.orc_mel_fbanks.argtypes = [c_int, c_double, c_double, c_int, c_int, POINTER(c_float)]
        L.orc_mel_fbanks_ex.argtypes = [c_int, c_double, c_double, c_int, c_int, c_int, POINTER(c_float)]
        L.orc_melspec_ex.argtypes = [POINTER(c_float), c_int, c_int, c_int, c_int, c_int, c_double, c_double, c_int,
                                     c_int, c_int, c_int, POINTER(c_float)]
        L.orc_sepconv_ex.argtypes = [POINTER(c_float)] + [c_int] * 5 + [POINTER(c_float)] * 8 + [c_int] * 5 + \
            [POINTER(c_float)] * 2
        L.orc_melspec.argtypes = [POINTER(c_float), c_int, c_int, c_int, c_int, c_int, c_double,
                                  c_double, c_int, POINTER(c_float)]
        L.orc_frame_pcm16.argtypes = [POINTER(c_int16), c_int64, c_int, c_int, POINTER(c_float), c_int64]
        L.orc_frame_pcm16.restype = c_int64
        L.orc_sepconv.argtypes = [POINTER(c_float), c_int, c_int, c_int, c_int, c_int] + \
            [POINTER(c_float)] * 8 + [c_int, POINTER(c_float), POINTER(c_float)]
        L.orc_head.argtypes = [POINTER(c_float), c_int, c_int, c_int, c_int] + [POINTER(c_float)] * 4 + \
            [c_int, POINTER(c_float)]
        L.orc_flat_ip_search.argtypes = [POINTER(c_float), c_int64, c_int, POINTER(c_float), c_int, c_int,
                                         POINTER(c_float), POINTER(c_int64)]
        L.orc_seq_score.argtypes = [POINTER(c_float), c_int, POINTER(c_int64), c_int, POINTER(c_float), c_int,
                                    POINTER(c_int64), c_int, POINTER(c_float), c_int, c_float]
        L.orc_seq_score.restype = c_int
        _LIB = L
    return _LIB


def ref_lib():
    """The reference's own seq_score (cpp/seqscore.cpp) built by oracle/Makefile, or None."""
    global _REF
    if _REF is None:
        so = os.path.join(_HERE, '_ref', 'seqscore.so')
        if not os.path.exists(so):
            build()
        if not os.path.exists(so):
            return None
        R = ctypes.CDLL(so)
        # prototypes exactly as database.py:16-29 declares them
        R.seq_score.argtypes = [c_void_p, POINTER(c_int64), c_int, POINTER(c_float), c_int,
                                POINTER(c_int64), c_int, POINTER(c_float), c_int, c_float]
        R.seq_score.restype = c_int
        R.version.restype = c_int64
        R.ref_flat_index_new.argtypes = [POINTER(c_float), c_int64, c_int]
        R.ref_flat_index_new.restype = c_void_p
        R.ref_flat_index_free.argtypes = [c_void_p]
        assert R.version() == 20220625002  # database.py:30
        _REF = R
    return _REF


# ---------------------------------------------------------------- stage 1
def mel_fbanks(params):
    n_freqs = params['stft_n'] // 2 + 1
    fb = np.empty((n_freqs, params['n_mels']), np.float32)
    lib().orc_mel_fbanks_ex(n_freqs, params['f_min'], params['f_max'], params['n_mels'],
                            params['sample_rate'], int(bool(params.get('naf_mode', False))), _p(fb, c_float))
    return fb


MEL_LOG = {None: 0, 'none': 0, 'log': 1, 'log10': 2}


def melspec(x, params):
    """MelSpec.forward (datautil/melspec.py:33-50) with the options build_mel_spec_layer reads from the config
    (melspec.py:60-62): naf_mode, mel_log ('log' | 'log10' | anything else = no logarithm), spec_norm ('l2' | 'max')."""
    x = _f32(x)
    B, n = x.shape
    T = 1 + n // params['stft_hop']
    out = np.empty((B, params['n_mels'], T), np.float32)
    rc = lib().orc_melspec_ex(_p(x, c_float), B, n, params['sample_rate'], params['stft_n'],
                              params['stft_hop'], params['f_min'], params['f_max'], params['n_mels'],
                              int(bool(params.get('naf_mode', False))), MEL_LOG.get(params.get('mel_log', 'log'), 0),
                              int(params.get('spec_norm', 'l2') == 'max'), _p(out, c_float))
    assert rc == 0, rc
    return out


def frame_pcm16(pcm, seg, hop):
    """datautil/musicdata.py:48,82-88 on mono int16 PCM at the target rate."""
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    n = pcm.shape[0]
    n_seg = (max(n, seg) - seg) // hop + 1
    rows = np.empty((n_seg, seg), np.float32)
    got = lib().orc_frame_pcm16(_p(pcm, c_int16), n, seg, hop, _p(rows, c_float), n_seg)
    assert got == n_seg
    return rows


# ---------------------------------------------------------------- stage 2
def layer_strides(params):
    """(time stride of conv1, frequency stride of conv2) per layer: model.py:82-85."""
    st = params['model'].get('strides')
    if st is None:
        return [(2, 2)] * 8
    return [(int(st[i][0][1]), int(st[i][1][0])) for i in range(8)]


def layer_shapes(params, F, T):
    """(Cin, Cout, F, T) of every SeparableConv2d, as MyF.__init__ builds them (model.py:79-93)."""
    m = params['model']
    d, h = m['d'], m['h']
    ch = [1, d, d, 2 * d, 2 * d, 4 * d, 4 * d, h, h]
    out = []
    for i, (st, sf) in enumerate(layer_strides(params)):
        out.append((ch[i], ch[i + 1], F, T))
        F = (F - 1) // sf + 1
        T = (T - 1) // st + 1
    assert F == 1 and T == 1, 'output must be 1x1'  # model.py:94
    return out


def fpnetwork_forward(sd, x, params, norm=True, return_layers=False):
    """FpNetwork.forward (model.py:148-153) from a reference-keyed state dict of numpy arrays."""
    L = lib()
    x = _f32(x)
    B, F, T = x.shape
    m = params['model']
    fuller = int(bool(m.get('fuller', False)))
    act = {'ReLU': 0, 'ELU': 1}[m.get('conv_activation', 'ReLU')]
    after = int(bool(m.get('relu_after_bn', True)))
    strides = layer_strides(params)
    cur = x.reshape(B, 1, F, T)
    layers = []
    for l, (ci, co, f, t) in enumerate(layer_shapes(params, F, T)):
        g = lambda k: _f32(sd['f.convs.%d.%s' % (l, k)])
        s_t, s_f = strides[l]
        f2, t2 = (f - 1) // s_f + 1, (t - 1) // s_t + 1
        y = np.empty((B, co, f2, t2), np.float32)
        mid = np.empty((B, co, f, t2), np.float32)
        arrs = [g('conv1.weight'), g('conv1.bias'), g('ln1.weight'), g('ln1.bias'),
                g('conv2.weight'), g('conv2.bias'), g('ln2.weight'), g('ln2.bias')]
        cur = _f32(cur)
        rc = L.orc_sepconv_ex(_p(cur, c_float), B, ci, co, f, t, *[_p(a, c_float) for a in arrs],
                              fuller, s_t, s_f, act, after, _p(y, c_float), _p(mid, c_float))
        assert rc == 0
        layers.append((mid, y))
        cur = y
    h, d, u = m['h'], m['d'], m['u']
    z = np.empty((B, d), np.float32)
    hx = _f32(cur.reshape(B, h))
    w = [_f32(sd['g.linear1.weight']), _f32(sd['g.linear1.bias']), _f32(sd['g.linear2.weight']),
         _f32(sd['g.linear2.bias'])]
    rc = L.orc_head(_p(hx, c_float), B, d, h, u, *[_p(a, c_float) for a in w], int(norm), _p(z, c_float))
    assert rc == 0
    return (z, layers) if return_layers else z


# ---------------------------------------------------------------- stage 3
def flat_ip_search(db, q, k):
    """faiss IndexFlatIP.search contract (database.py:121)."""
    db, q = _f32(db), _f32(q)
    Q = q.shape[0]
    dist = np.empty((Q, k), np.float32)
    lab = np.empty((Q, k), np.int64)
    lib().orc_flat_ip_search(_p(db, c_float), db.shape[0], db.shape[1] if db.ndim == 2 else q.shape[1],
                             _p(q, c_float), Q, k, _p(dist, c_float), _p(lab, c_int64))
    return dist, lab


def seq_score(db, song_pos, query, labels, frame_shift_mul=1, score_alpha=0.0, use_ref=False):
    """cpp/seqscore.cpp:33-136.  Returns (best_song, song_scores[n_songs,2]) exactly as the C
    function leaves them (times in frames, not yet rescaled by database.py:191-193)."""
    db, query = _f32(db), _f32(query)
    song_pos = np.ascontiguousarray(song_pos, dtype=np.int64)
    labels = np.ascontiguousarray(labels, dtype=np.int64)
    n_songs = song_pos.shape[0] - 1
    ss = np.zeros((n_songs, 2), np.float32)  # database.py:176
    if use_ref:
        R = ref_lib()
        assert R is not None, 'oracle/_ref/seqscore.so missing'
        ix = R.ref_flat_index_new(_p(db, c_float), db.shape[0], query.shape[1])
        try:
            best = R.seq_score(ix, _p(song_pos, c_int64), n_songs, _p(query, c_float), query.shape[0],
                               _p(labels, c_int64), labels.shape[1], _p(ss, c_float), frame_shift_mul,
                               score_alpha)
        finally:
            R.ref_flat_index_free(ix)
    else:
        best = lib().orc_seq_score(_p(db, c_float), query.shape[1], _p(song_pos, c_int64), n_songs,
                                   _p(query, c_float), query.shape[0], _p(labels, c_int64), labels.shape[1],
                                   _p(ss, c_float), frame_shift_mul, score_alpha)
    return best, ss


def query_embeddings_cpp(db, song_pos, query, labels, frame_shift_mul, hop_size, score_alpha=0.0,
                         use_ref=False):
    """Database.query_embeddings_cpp post-processing (database.py:190-195)."""
    song_id, ss = seq_score(db, song_pos, query, labels, frame_shift_mul, score_alpha, use_ref)
    best = ss[song_id, 0].item()
    best_song_t = song_id, ss[song_id, 1].item() * hop_size / frame_shift_mul
    ss[:, 1] *= hop_size / frame_shift_mul
    return best, best_song_t, ss


def query_embeddings_base(db, song_pos, query, labels, frame_shift_mul, hop_size):
    """numpy restatement of Database.query_embeddings_base (database.py:117-166) with
    index.reconstruct(i) == db[i].  Small cases only (Python loops, like the reference)."""
    db, query = _f32(db), _f32(query)
    song_pos = np.asarray(song_pos, dtype=np.int64)
    n_songs = song_pos.shape[0] - 1
    best = -1e999
    best_song_t = -1, 0
    song_score = np.zeros([n_songs, 2], dtype=np.float32)
    if db.shape[0] == 0:
        return best, best_song_t, song_score
    for shift in range(frame_shift_mul):
        cands = []
        sub = query[shift::frame_shift_mul]
        sub_len = sub.shape[0]
        for t in range(sub_len):
            lab = labels[t * frame_shift_mul + shift]
            lab = lab[lab != -1]
            sid = np.searchsorted(song_pos, lab, side='right') - 1
            cands.append(np.stack([sid, lab - song_pos[sid] - t], axis=1))
        cands = np.unique(np.concatenate(cands), axis=0)
        vec = np.zeros_like(sub)
        for c in cands:
            sid = int(c[0])
            start = int(song_pos[sid])
            slen = int(song_pos[sid + 1]) - start
            t = int(c[1])
            real_time = (t - shift / frame_shift_mul) * hop_size
            for i in range(sub_len):
                if t + i < 0 or t + i >= slen:
                    vec[i] = 0.0
                else:
                    vec[i] = db[start + t + i]
            sco = np.dot(vec.flatten(), sub.flatten()).item() / sub_len
            if sco > song_score[sid, 0]:
                song_score[sid, 0] = sco
                song_score[sid, 1] = real_time
            if sco > best:
                best = sco
                best_song_t = sid, real_time
    return best, best_song_t, song_score


# ---------------------------------------------------------------- training-step pieces
def similarity_loss(y, tau):
    """train.py:41-52 in float64, with its analytic gradient: (loss, dL/dy)."""
    y = np.asarray(y, np.float64)
    N = y.shape[0]
    a = y @ y.T / tau
    np.fill_diagonal(a, -np.inf)                      # the row's own entry is left out (train.py:46)
    m = a.max(axis=1, keepdims=True)
    lse = m[:, 0] + np.log(np.exp(a - m).sum(axis=1))
    p = np.arange(N) ^ 1                              # train.py:48: partner of row i
    loss = -(a[np.arange(N), p] - lse).sum() / N
    P = np.exp(a - lse[:, None])
    G = P.copy()
    G[np.arange(N), p] -= 1.0
    dy = (G + G.T) @ y / (tau * N)
    return loss, dy


def add_noises(x, noise, snr_db):
    """datautil/noise.py:96-109 given the noise rows and the SNRs."""
    x, noise = np.asarray(x, np.float64), np.asarray(noise, np.float64)
    vx = np.sqrt(np.maximum((x ** 2).mean(axis=1), 1e-12))
    vn = np.sqrt(np.maximum((noise ** 2).mean(axis=1), 1e-12))
    ratio = vx / vn * 10.0 ** (-np.asarray(snr_db, np.float64) / 20.0)
    return x + ratio[:, None] * noise


# ---------------------------------------------------------------- ingest
def resample_frac(x, old_sr, new_sr, zeros=24, rolloff=0.945):
    """julius.ResampleFrac(old_sr, new_sr)(x) for x[..., n] -- julius is a third-party dependency of the reference
    (musicdata.py:1,29; unpinned, readme.md:20) that is absent here, so this restates its published algorithm and is
    NOT pinned to julius itself ("parity unpinned"): windowed-sinc kernels per output phase, replicate padding,
    output length int(new_sr * n / old_sr)."""
    import math
    x = np.asarray(x, np.float64)
    g = math.gcd(old_sr, new_sr)
    o, w = old_sr // g, new_sr // g
    if o == w:
        return x.copy()
    sr = min(o, w) * rolloff
    width = math.ceil(zeros * o / sr)
    idx = np.arange(-width, width + o, dtype=np.float64)
    kern = []
    for i in range(w):
        t = np.clip((-i / w + idx / o) * sr, -zeros, zeros) * math.pi
        k = np.sinc(t / math.pi) * np.cos(t / zeros / 2) ** 2
        kern.append(k / k.sum())
    kern = np.stack(kern)                                          # [new, K]
    n = x.shape[-1]
    xp = np.concatenate([np.repeat(x[..., :1], width, -1), x, np.repeat(x[..., -1:], width + o, -1)], -1)
    n_frames = (xp.shape[-1] - kern.shape[1]) // o + 1
    win = np.lib.stride_tricks.sliding_window_view(xp, kern.shape[1], axis=-1)[..., ::o, :][..., :n_frames, :]
    y = np.einsum('...mk,ik->...mi', win, kern).reshape(*x.shape[:-1], -1)
    return y[..., :int(w * n / o)]


def mix_mono(wav):
    """musicdata.py:72-80: wav[nch, n] -> mono[n] with the fake-stereo rule (fp32 like the reference)."""
    wav = np.array(wav, np.float32)
    if wav.shape[0] == 2:
        pow1 = ((wav[0] - wav[1]) ** 2).mean()
        pow2 = ((wav[0] + wav[1]) ** 2).mean()
        if pow1 > pow2 * 1000:
            wav[1] *= -1
    return wav.mean(axis=0, dtype=np.float32)


def apply_ir(x, responses, pad_start=0, segment_size=None):
    """datautil/dataset_v2.py:157-163 in float64, direct form: rows of x convolved with their impulse responses one after
    the other (the reference multiplies spectra of an FFT long enough that nothing wraps: a causal linear convolution),
    samples [pad_start, segment_size) kept."""
    x = np.asarray(x, np.float64)
    end = x.shape[1] if segment_size is None else int(segment_size)
    out = np.empty((x.shape[0], end - pad_start))
    for b in range(x.shape[0]):
        y = x[b]
        for h in responses:
            y = np.convolve(y, np.asarray(h[b], np.float64))[:end]
        out[b] = y[pad_start:end]
    return out


def ln_act_backward(dA, Y, gamma, beta, act='ReLU', act_first=False, eps=1e-5):
    """Backward of one `LayerNorm over (C,F,T) + activation` block of SeparableConv2d (model.py:58-72) in float64, the
    formulas pfann_b200/csrc/encoder_train.cu implements: given dL/d(output) for Y[B, ...] (raw convolution output),
    returns (dL/dY, dL/dgamma, dL/dbeta).  act_first = relu_after_bn False: out = LN(act(Y)); else out = act(LN(Y))."""
    Y = np.asarray(Y, np.float64)
    dA = np.asarray(dA, np.float64)
    B = Y.shape[0]
    red = tuple(range(1, Y.ndim))

    def act_f(v):
        return np.where(v > 0, v, np.expm1(np.minimum(v, 0))) if act == 'ELU' else np.maximum(v, 0)

    def act_d(v):
        return np.where(v > 0, 1.0, np.exp(np.minimum(v, 0))) if act == 'ELU' else (v > 0).astype(np.float64)

    z = act_f(Y) if act_first else Y
    mean = z.mean(axis=red, keepdims=True)
    rstd = 1.0 / np.sqrt(z.var(axis=red, keepdims=True) + eps)
    xh = (z - mean) * rstd
    n = xh * gamma + beta
    dn = dA if act_first else dA * act_d(n)
    g = dn * gamma
    dz = rstd * (g - g.mean(axis=red, keepdims=True) - xh * (g * xh).mean(axis=red, keepdims=True))
    dY = dz * act_d(Y) if act_first else dz
    return dY, (dn * xh).sum(axis=0), dn.sum(axis=0)
