#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference code (/root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python tools/gen_golden.py

The reference ships no golden vectors, so parity is pinned to outputs of
the reference's own Python modules executed here:

  * datautil/melspec.py  (torchaudio present)             -> mel_default.npz
  * model.py FpNetwork   (load_state_dict of our seeded weights) -> enc_<config>.npz
  * database.py Database.query_embeddings_base, imported unmodified on top of a ~40-line numpy
    `faiss` shim (faiss is not installed and cannot be: no network)   -> db_small.npz
  * datautil/musicdata.py MusicDataset, unmodified, over an identity `julius` shim (the
    sources are already at 8 kHz, where julius.ResampleFrac is the identity) -> musicdata.npz

Inputs are regenerated from seeds by pfann_b200.synth, so only outputs are stored.
"""
import os
import sys
import tempfile
import types
import wave

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('PFANN_REFERENCE', '/root/reference')
sys.path.insert(0, REPO)
sys.path.insert(0, REF)
from pfann_b200 import synth  # noqa: E402

OUT = os.path.join(REPO, 'tests', 'golden')


# ----------------------------------------------------------------------------- shims
class _FlatIP:
    """IndexFlatIP contract: exact fp32 inner product, descending, ties -> lower id, -1 pad."""

    def __init__(self, d):
        self.d = d
        self.x = np.zeros((0, d), np.float32)
        self.is_trained = True

    @property
    def ntotal(self):
        return self.x.shape[0]

    def add(self, x):
        self.x = np.concatenate([self.x, np.asarray(x, np.float32)])

    def reconstruct(self, i, out=None):
        if out is None:
            return self.x[i].copy()
        out[:] = self.x[i]
        return out

    def search(self, q, k):
        s = np.asarray(q, np.float32) @ self.x.T
        n = self.ntotal
        order = np.lexsort((np.broadcast_to(np.arange(n), s.shape), -s), axis=1)[:, :k]
        D = np.full((q.shape[0], k), -np.finfo(np.float32).max, np.float32)
        I = np.full((q.shape[0], k), -1, np.int64)
        D[:, :order.shape[1]] = np.take_along_axis(s, order, 1)
        I[:, :order.shape[1]] = order
        return D, I


def install_shims():
    faiss = types.ModuleType('faiss')
    faiss.IndexFlatIP = _FlatIP
    faiss.IndexFlat = _FlatIP
    for n in ('Index', 'IndexBinary', 'IndexPreTransform', 'IndexIVF'):
        setattr(faiss, n, type(n, (), {}))
    faiss.METRIC_INNER_PRODUCT, faiss.METRIC_L2 = 0, 1
    faiss.downcast_index = lambda ix: ix
    store = {}
    faiss.write_index = lambda ix, path: store.__setitem__(path, ix)
    faiss.read_index = lambda path: store[path]
    faiss._store = store
    sys.modules['faiss'] = faiss
    julius = types.ModuleType('julius')

    class ResampleFrac:
        def __init__(self, old_sr, new_sr):
            assert old_sr == new_sr, 'shim is the identity resampler only'

        def __call__(self, x):
            return x
    julius.ResampleFrac = ResampleFrac
    sys.modules['julius'] = julius
    return faiss


# ----------------------------------------------------------------------------- fixtures
def golden_mel(params):
    from datautil.melspec import build_mel_spec_layer
    x = np.concatenate([synth.synth_segments(3, seed=1), np.zeros((1, 8000), np.float32)])
    mel = build_mel_spec_layer(params).eval()
    with torch.no_grad():
        y = mel(torch.from_numpy(x)).numpy()
    np.savez_compressed(os.path.join(OUT, 'mel_default.npz'), mel=y.astype(np.float32),
                        note='x = concat(synth_segments(3, seed=1), zeros(1,8000))')
    return y


# option sets of MelSpec beyond the default (melspec.py:27-49): NAF-converted models use naf_mode + log10 + max
MEL_VARIANTS = {
    'naf': {'naf_mode': True, 'mel_log': 'log10', 'spec_norm': 'max'},
    'log10': {'mel_log': 'log10'},
    'max': {'spec_norm': 'max'},
    'nolog_naf': {'naf_mode': True, 'mel_log': 'none'},
}
# option sets of FpNetwork beyond the default (model.py:58-72,84-85), on the tiny config (d=8, h=32) and a
# NAF-style stride schedule for F=256, T=32
MODEL_VARIANTS = {
    'elu': {'conv_activation': 'ELU'},
    'act_first': {'relu_after_bn': False},
    'elu_act_first': {'conv_activation': 'ELU', 'relu_after_bn': False},
    'strides': {'strides': [[[1, 2], [2, 1]], [[1, 2], [2, 1]], [[1, 2], [2, 1]], [[1, 2], [2, 1]],
                            [[1, 1], [2, 1]], [[1, 2], [2, 1]], [[1, 1], [2, 1]], [[1, 1], [2, 1]]]},
}


def golden_mel_variants(params):
    from datautil.melspec import build_mel_spec_layer
    x = np.concatenate([synth.synth_segments(3, seed=1), np.zeros((1, 8000), np.float32)])
    out = {}
    for name, opt in MEL_VARIANTS.items():
        p = dict(params, **opt)
        mel = build_mel_spec_layer(p).eval()
        with torch.no_grad():
            out[name] = mel(torch.from_numpy(x)).numpy().astype(np.float32)
    np.savez_compressed(os.path.join(OUT, 'mel_variants.npz'), **out)


def golden_model_variants(mel):
    from model import FpNetwork
    base = synth.read_config('tiny')
    out = {}
    for name, opt in MODEL_VARIANTS.items():
        params = dict(base, model=dict(base['model'], **opt))
        d, h, u, F, T = synth.model_dims(params)
        sd = synth.make_state_dict(params, seed=21)
        net = FpNetwork(d, h, u, F, T, params['model']).eval()
        net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        with torch.no_grad():
            x = torch.from_numpy(mel)
            out['z_' + name] = net(x, norm=True).numpy()
            cur = x.unsqueeze(1)
            for i, conv in enumerate(net.f.convs):
                cur = conv(cur)
                if i == 0:
                    out['l0_' + name] = cur.numpy().astype(np.float32)
            out['enc_' + name] = cur.reshape(cur.shape[0], -1).numpy().astype(np.float32)
    np.savez_compressed(os.path.join(OUT, 'enc_variants.npz'), **out)


def _reference_function(path, name):
    """The named top-level function of a reference source file, executed as written (train.py imports tensorboardX and
    torch_optimizer at module level, which are not installed: the function itself only needs torch)."""
    import ast
    src = open(os.path.join(REF, path)).read()
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            ns = {'torch': torch, 'np': np}
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, 'exec'), ns)
            return ns[name]
    raise KeyError(name)


def golden_train():
    """similarity_loss of train.py:41-52 (value + autograd gradient) and SpecAugment masks of specaug.py:13-38."""
    loss_fn = _reference_function('train.py', 'similarity_loss')
    out = {}
    for tag, (N, d, seed) in {'n640d64': (640, 64, 5), 'n8d16': (8, 16, 6)}.items():
        rng = np.random.Generator(np.random.PCG64(seed))
        y = rng.standard_normal((N, d)).astype(np.float32)
        y[1::2] = y[0::2] + 0.5 * y[1::2]                       # positive pairs are correlated
        y /= np.linalg.norm(y, axis=1, keepdims=True)
        for dt, sfx in ((torch.float32, ''), (torch.float64, '_f64')):
            t = torch.from_numpy(y).to(dt).requires_grad_(True)
            loss = loss_fn(t, 0.05)
            loss.backward()
            out['loss_%s%s' % (tag, sfx)] = np.float64(loss.item())
            out['dy_%s%s' % (tag, sfx)] = t.grad.numpy().astype(np.float32)
        out['seed_' + tag] = seed
    from datautil.specaug import SpecAugment
    sa = SpecAugment({'cutout_min': 0.1, 'cutout_max': 0.5})
    torch.manual_seed(1234)
    out['specaug_masks'] = np.stack([sa.get_mask(256, 32).numpy() for _ in range(6)]).astype(np.uint8)
    torch.manual_seed(77)
    x = torch.arange(2 * 256 * 32, dtype=torch.float32).reshape(2, 256, 32) + 1
    out['specaug_x'] = sa.augment(x).numpy()
    # impulse responses: the statements of datautil/dataset_v2.py:51-58,157-163 and ir.py:37-38,72-73 executed as
    # written (the classes around them need the AIR / microphone data sets, which are not available offline)
    rng = np.random.Generator(np.random.PCG64(41))
    pad_start, segment_size_total = 100, 8100                       # pad_start + segment_size (dataset_v2.py:133)
    x_aug = torch.from_numpy(synth.synth_segments(3, seed=8, seg=segment_size_total))
    air = (rng.standard_normal((3, 8000)) * np.exp(-np.arange(8000) / 900.0)).astype(np.float32)
    mic = (rng.standard_normal((3, 4000)) * np.exp(-np.arange(4000) / 60.0)).astype(np.float32)
    fftconv_n = 1024
    while fftconv_n < segment_size_total + 8000 + 4000:
        fftconv_n *= 2
    spec = torch.fft.rfft(x_aug, fftconv_n)
    spec *= torch.fft.rfft(torch.from_numpy(air), fftconv_n)
    spec *= torch.fft.rfft(torch.from_numpy(mic), fftconv_n)
    y = torch.fft.irfft(spec, fftconv_n)[..., pad_start:segment_size_total]
    out['ir_air'], out['ir_mic'], out['ir_out'] = air, mic, y.numpy().astype(np.float32)
    np.savez_compressed(os.path.join(OUT, 'train.npz'), **out)


def golden_train_step():
    """One training step of the reference network under torch autograd (train.py:96-103 without the optimizer):
    z = model(x), loss = similarity_loss(z, tau), loss.backward().  Per parameter the fixture keeps the gradient's
    l2 norm and a seeded sample of its elements."""
    from model import FpNetwork
    loss_fn = _reference_function('train.py', 'similarity_loss')
    out = {}
    for name, (cfg, opt, B, seed) in synth.TRAIN_CASES.items():
        base = synth.read_config(cfg)
        params = dict(base, model=dict(base['model'], **opt))
        d, h, u, F, T = synth.model_dims(params)
        sd = synth.make_state_dict(params, seed=seed)
        net = FpNetwork(d, h, u, F, T, params['model']).train()
        net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        x = torch.from_numpy(synth.train_case_input(name))
        z = net(x)
        z.retain_grad()
        loss = loss_fn(z, params['tau'])
        loss.backward()
        out[name + '/z'] = z.detach().numpy()
        out[name + '/dz'] = z.grad.numpy()
        out[name + '/loss'] = np.float64(loss.item())
        for i, (k, p) in enumerate(net.named_parameters()):
            g = p.grad.numpy().reshape(-1)
            out['%s/norm/%s' % (name, k)] = np.float64(np.sqrt((g.astype(np.float64) ** 2).sum()))
            out['%s/vals/%s' % (name, k)] = g[synth.grad_sample_index(g.size, seed * 100 + i)].astype(np.float32)
    np.savez_compressed(os.path.join(OUT, 'train_step.npz'), **out)


def golden_encoder(name, params, mel, seed):
    from model import FpNetwork
    d, h, u, F, T = synth.model_dims(params)
    sd = synth.make_state_dict(params, seed=seed)
    net = FpNetwork(d, h, u, F, T, params['model']).eval()
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    x = torch.from_numpy(mel)
    acts = []
    with torch.no_grad():
        z = net(x, norm=True).numpy()
        zr = net(x, norm=False).numpy()
        cur = x.unsqueeze(1)
        for conv in net.f.convs:
            cur = conv(cur)
            acts.append([cur.mean().item(), cur.std().item(), cur.abs().max().item()])
        enc = cur.reshape(cur.shape[0], -1).numpy()
    np.savez_compressed(os.path.join(OUT, 'enc_%s.npz' % name), z=z, z_raw=zr, layer_stats=np.array(acts),
                        enc_out=enc.astype(np.float32), seed=seed)


def golden_db(faiss):
    import database
    rng = np.random.Generator(np.random.PCG64(42))
    d = 16
    key = np.array([30, 0, 5, 59, 19, 40, 1, 25, 59, 33, 12, 59], np.int32)  # includes empty + tiny songs
    n = int(key.sum())
    db = rng.standard_normal((n, d), dtype=np.float32)
    db /= np.linalg.norm(db, axis=1, keepdims=True)
    pos = synth.song_pos_from_key(key)
    cases = []
    tmp = tempfile.mkdtemp()
    with open(os.path.join(tmp, 'songList.txt'), 'w') as f:
        f.write('\n'.join('song%d.wav' % i for i in range(len(key))) + '\n')
    key.tofile(os.path.join(tmp, 'landmarkKey'))
    ix = _FlatIP(d)
    ix.add(db)
    faiss._store[os.path.join(tmp, 'landmarkValue')] = ix
    out = {'db': db, 'key': key}
    ci = 0
    for fsm in (1, 2, 4):
        for (song, off, qlen) in ((3, 10, 19), (0, -3, 19), (11, 50, 19), (4, 0, 7), (7, 20, 38 if fsm > 1 else 19)):
            # a noisy copy of a database diagonal; may overhang either end of the song
            rows = []
            for j in range(qlen):
                t = off + j // fsm
                if 0 <= t < key[song]:
                    v = db[pos[song] + t].copy()
                else:
                    v = rng.standard_normal(d, dtype=np.float32)
                v = v + rng.standard_normal(d, dtype=np.float32) * np.float32(0.25 / np.sqrt(d) * 4)
                rows.append(v / np.linalg.norm(v))
            q = np.stack(rows).astype(np.float32)
            dbo = database.Database(tmp, {'top_k': 8, 'frame_shift_mul': fsm}, 0.5)
            _, labels = ix.search(q, 8)
            sco, (sid, tim), ss = dbo.query_embeddings(q)
            out.update({'q%d' % ci: q, 'labels%d' % ci: labels, 'fsm%d' % ci: fsm, 'score%d' % ci: np.float64(sco),
                        'song%d' % ci: sid, 'time%d' % ci: np.float64(tim), 'ss%d' % ci: ss})
            cases.append(ci)
            ci += 1
    out['n_cases'] = ci
    np.savez_compressed(os.path.join(OUT, 'db_small.npz'), **out)


def golden_musicdata(params):
    from datautil.musicdata import MusicDataset
    tmp = tempfile.mkdtemp()
    lens = [8000, 20000, 5000, 80000]
    paths = []
    for i, n in enumerate(lens):
        p = os.path.join(tmp, 'clip%d.wav' % i)
        with wave.open(p, 'wb') as w:
            w.setnchannels(1)
            w.setsampwidth(2)
            w.setframerate(8000)
            w.writeframes(synth.synth_pcm(500 + i, n).tobytes())
        paths.append(p)
    lst = os.path.join(tmp, 'list.txt')
    with open(lst, 'w') as f:
        f.write('\n'.join(paths) + '\n')
    out = {'lens': np.array(lens)}
    for fsm in (1, 2):
        p = dict(params)
        p['indexer'] = dict(params['indexer'], frame_shift_mul=fsm)
        ds = MusicDataset(lst, p)
        for i in range(len(lens)):
            _, _, wav = ds[i]
            w = wav.numpy()
            out['nseg_f%d_c%d' % (fsm, i)] = w.shape[0]
            if lens[i] <= 20000:
                out['rows_f%d_c%d' % (fsm, i)] = w.astype(np.float32)
            else:  # keep the fixture small: row means / norms + first row
                out['rownorm_f%d_c%d' % (fsm, i)] = np.linalg.norm(w.astype(np.float64), axis=1)
                out['row0_f%d_c%d' % (fsm, i)] = w[0].astype(np.float32)
    np.savez_compressed(os.path.join(OUT, 'musicdata.npz'), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(8)
    faiss = install_shims()
    import multiprocessing as mp
    mp.get_logger()
    default = synth.read_config('default')
    mel = golden_mel(default)
    golden_encoder('default', default, mel, seed=11)
    golden_encoder('n640d64', synth.read_config('n640d64'), mel, seed=12)
    golden_encoder('tiny', synth.read_config('tiny'), mel, seed=13)
    golden_mel_variants(default)
    golden_model_variants(mel)
    golden_train()
    golden_train_step()
    golden_db(faiss)
    golden_musicdata(default)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == '__main__':
    main()
