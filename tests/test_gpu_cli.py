"""GPU (-m gpu): BASELINE config 1 -- the drop-in command lines end to end on 10 synthetic 1 s WAVs:
builder.py writes the reference's database layout, matcher.py finds every file as its own best match at
offset 0, and the extractemb.py + matchemb.py split gives the same answers."""
import csv
import json
import os
import sys
import wave

import numpy as np
import pytest

torch = pytest.importorskip('torch')
from pfann_b200 import synth  # noqa: E402

pytestmark = pytest.mark.gpu


def _wav(path, pcm, nch=1):
    with wave.open(path, 'wb') as w:
        w.setnchannels(nch)
        w.setsampwidth(2)
        w.setframerate(8000)
        w.writeframes(pcm.tobytes())


def test_builder_matcher_roundtrip(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from pfann_b200 import cli
    d = str(tmp_path)
    params = synth.read_config('n640d64')
    params['indexer'] = {'index_factory': 'Flat', 'top_k': 5, 'frame_shift_mul': 1}
    model_dir = os.path.join(d, 'model')
    os.makedirs(model_dir)
    params['model_dir'] = model_dir
    json.dump(params, open(os.path.join(model_dir, 'configs.json'), 'w'))
    sd = synth.make_state_dict(params, seed=3)
    torch.save({k: torch.from_numpy(v) for k, v in sd.items()}, os.path.join(model_dir, 'model.pt'))
    files = []
    for i in range(10):                       # 0.5 sin + 0.1 noise, seed 1000+i
        rng = np.random.Generator(np.random.PCG64(1000 + i))
        n = 8000 if i < 8 else (20000 if i == 8 else 5000)       # plus one long and one short clip
        t = np.arange(n) / 8000.0
        x = 0.5 * np.sin(2 * np.pi * (400 + 300 * i) * t) + 0.1 * rng.standard_normal(n)
        p = os.path.join(d, 'clip%d.wav' % i)
        _wav(p, np.round(np.clip(x, -1, 1) * 32767).astype(np.int16))
        files.append(p)
    files.append(os.path.join(d, 'missing.wav'))                 # unreadable -> 0 segments, no abort
    lst = os.path.join(d, 'list.txt')
    open(lst, 'w').write('\n'.join(files) + '\n')
    db_dir = os.path.join(d, 'db')
    assert cli.builder_main(['builder.py', lst, db_dir, model_dir]) == 0
    for f in ('embeddings', 'landmarkValue', 'landmarkKey', 'songList.txt', 'configs.json', 'model.pt'):
        assert os.path.exists(os.path.join(db_dir, f)), f
    key = np.fromfile(os.path.join(db_dir, 'landmarkKey'), dtype=np.int32)
    assert list(key) == [1] * 8 + [4, 1, 0]
    emb = np.fromfile(os.path.join(db_dir, 'embeddings'), dtype=np.float32).reshape(-1, 64)
    assert emb.shape[0] == key.sum()
    np.testing.assert_allclose(np.linalg.norm(emb, axis=1), 1.0, atol=1e-5)
    res = os.path.join(d, 'result.txt')
    assert cli.matcher_main(['matcher.py', lst, db_dir, res]) == 0
    lines = [l.rstrip('\n').split('\t') for l in open(res, encoding='utf8')]
    assert len(lines) == 11
    for i in range(10):
        assert lines[i] == [files[i], files[i]]
    assert lines[10][1] == 'error'
    rows = list(csv.reader(open(os.path.join(d, 'result_detail.csv'))))
    assert rows[0] == ['query', 'answer', 'score', 'time', 'part_scores']
    for i in range(10):
        assert rows[1 + i][1] == files[i] and float(rows[1 + i][3]) == 0.0 and float(rows[1 + i][2]) > 0.99
    binf = np.fromfile(res + '.bin', dtype=np.float32).reshape(11, 11, 2)
    assert (binf[10] == 0).all() and all(binf[i, i, 0] > 0.99 for i in range(10))
    # the split pipeline (extractemb.py -> matchemb.py) gives the same answers
    qdir = os.path.join(d, 'qemb')
    assert cli.extractemb_main(['extractemb.py', lst, db_dir, qdir]) == 0
    qi = np.fromfile(os.path.join(qdir, 'query_index'), dtype=np.int64).reshape(-1, 2)
    assert list(qi[:, 1]) == list(key)
    res2 = os.path.join(d, 'result2.txt')
    assert cli.matchemb_main(['matchemb.py', qdir, db_dir, res2]) == 0
    assert open(res2).read() == open(res).read()
    assert np.array_equal(np.fromfile(res2 + '.bin', dtype=np.float32), np.fromfile(res + '.bin', dtype=np.float32))
