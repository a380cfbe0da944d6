"""CPU: the C-ABI library loads, exports every symbol include/pfann_b200.h declares, keeps the
reference-compatible handshake, and fails loudly (no CPU fallback) when no GPU is present."""
import ctypes
import os
import re

import numpy as np
import pytest

from pfann_b200 import _lib, synth

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(REPO, 'include', 'pfann_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    names = re.findall(r'\b([a-z_0-9]+)\s*\(', src)
    return sorted(set(n for n in names if n.startswith('pfann_') or n in ('version', 'seq_score')))


def test_library_builds_and_exports_every_declared_symbol():
    L = _lib.lib()
    syms = _declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(L, s), 'libpfann_b200.so does not export %s' % s


def test_version_handshake():
    L = _lib.lib()
    assert L.pfann_version() == _lib.VERSION
    assert L.version() == 20220625002          # cpp/seqscore.cpp:27-30, checked like database.py:30


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    h = ctypes.c_void_p()
    rc = _lib.lib().pfann_ctx_create(0, ctypes.byref(h))
    assert rc < 0
    assert b'no CPU fallback' in _lib.lib().pfann_last_error()
    with pytest.raises(_lib.PfannError):
        _lib.ctx(0)
    # the Python mirrors refuse too instead of computing on the CPU
    from pfann_b200.datautil.melspec import build_mel_spec_layer
    mel = build_mel_spec_layer(synth.read_config('default'))
    with pytest.raises(_lib.PfannError):
        mel(torch.zeros(1, 8000))


def test_count_segments_matches_musicdata_rule():
    L = _lib.lib()
    lens = [8000, 20000, 5000, 80000, 0, 7999, 8001, 12000]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    for hop in (4000, 2000, 1000):
        want = sum((max(n, 8000) - 8000) // hop + 1 for n in lens)      # musicdata.py:82-87
        got = L.pfann_count_segments(off.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), len(lens), 8000, hop)
        assert got == want


def test_model_mirror_has_reference_state_dict_keys():
    from pfann_b200.model import FpNetwork
    for name in ('default', 'n640d64', 'tiny'):
        params = synth.read_config(name)
        d, h, u, F, T = synth.model_dims(params)
        net = FpNetwork(d, h, u, F, T, params['model'])
        sd = synth.make_state_dict(params, seed=1)
        got = {k: tuple(v.shape) for k, v in net.state_dict().items()}
        want = {k: tuple(v.shape) for k, v in sd.items()}
        assert got == want
        import torch
        net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})


def test_option_variants_construct_like_the_reference():
    """Every option of melspec.py:4-31 / model.py:132-146 is accepted (they have kernels now); the state_dict of a
    custom-stride network has the reference's shapes; an unknown activation fails like model.py:12."""
    from pfann_b200.datautil.melspec import MelSpec
    from pfann_b200.model import FpNetwork
    MelSpec(naf_mode=True, mel_log='log10', spec_norm='max')
    FpNetwork(128, 1024, 32, 256, 32, {'conv_activation': 'ELU', 'relu_after_bn': False})
    strides = [[[1, 2], [2, 1]]] * 4 + [[[1, 1], [2, 1]], [[1, 2], [2, 1]], [[1, 1], [2, 1]], [[1, 1], [2, 1]]]
    params = dict(synth.read_config('tiny'))
    params['model'] = dict(params['model'], strides=strides)
    d, h, u, F, T = synth.model_dims(params)
    net = FpNetwork(d, h, u, F, T, params['model'])
    sd = synth.make_state_dict(params, seed=1)
    assert {k: tuple(v.shape) for k, v in net.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}
    assert tuple(net.f.convs[4].ln1.normalized_shape) == (4 * d, 16, 2)       # time stride 1 in layer 4
    with pytest.raises(KeyError):
        FpNetwork(8, 32, 4, 256, 32, {'conv_activation': 'GELU'})


def test_flat_index_file_roundtrip(tmp_path):
    from pfann_b200.database import read_flat_index, write_flat_ip_index
    emb, _ = synth.synth_db(37, d=16, seed=2)
    p = str(tmp_path / 'landmarkValue')
    write_flat_ip_index(p, emb)
    back = read_flat_index(p)
    assert np.array_equal(back, emb)
    with open(p, 'rb') as f:
        assert f.read(4) == b'IxFI'
