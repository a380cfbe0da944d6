"""CPU: the oracle (oracle/) against the golden vectors produced by the reference's own code
(tools/gen_golden.py) and against the reference's own seqscore.cpp built into oracle/_ref/."""
import os

import numpy as np
import pytest

from oracle import pfann_oracle as orc
from pfann_b200 import synth


def _mel_inputs():
    return np.concatenate([synth.synth_segments(3, seed=1), np.zeros((1, 8000), np.float32)])


def test_mel_fbanks_structure():
    fb = orc.mel_fbanks(synth.read_config('default'))
    assert fb.shape == (513, 256)
    assert (fb[:39] == 0).all()                      # rows < 39 are all zero
    nnz = (fb > 0).sum(0)
    assert nnz.min() >= 1 and nnz.max() <= 7


def test_melspec_vs_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'mel_default.npz'))['mel']
    y = orc.melspec(_mel_inputs(), synth.read_config('default'))
    assert y.shape == g.shape == (4, 256, 32)
    # all-zero row: log(0 + 1e-8) everywhere (melspec.py:41,46)
    np.testing.assert_allclose(y[3], np.log(np.float32(1e-8)), rtol=0, atol=1e-6)
    np.testing.assert_allclose(g[3], np.log(np.float32(1e-8)), rtol=0, atol=1e-6)
    # tolerance: the reference computes in fp32 (cuFFT/pocketfft + sgemm); the oracle in double.
    err = np.abs(y[:3] - g[:3])
    assert err.max() < 2e-3, err.max()
    assert err.mean() < 2e-5, err.mean()


@pytest.mark.parametrize('name', ['tiny', 'n640d64', 'default'])
def test_encoder_vs_reference_golden(golden_dir, name):
    params = synth.read_config(name)
    g = np.load(os.path.join(golden_dir, 'enc_%s.npz' % name))
    mel = np.load(os.path.join(golden_dir, 'mel_default.npz'))['mel']
    n = 4 if name == 'tiny' else 2                   # keep the CPU suite short
    sd = synth.make_state_dict(params, seed=int(g['seed']))
    z = orc.fpnetwork_forward(sd, mel[:n], params, norm=True)
    zr, layers = orc.fpnetwork_forward(sd, mel[:n], params, norm=False, return_layers=True)
    np.testing.assert_allclose(z, g['z'][:n], rtol=0, atol=2e-5)
    np.testing.assert_allclose(zr, g['z_raw'][:n], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(layers[-1][1].reshape(n, -1), g['enc_out'][:n], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(np.linalg.norm(z, axis=1), 1.0, atol=1e-6)


MEL_VARIANTS = {     # tools/gen_golden.py MEL_VARIANTS: option sets of melspec.py:27-49 beyond the default
    'naf': {'naf_mode': True, 'mel_log': 'log10', 'spec_norm': 'max'},
    'log10': {'mel_log': 'log10'},
    'max': {'spec_norm': 'max'},
    'nolog_naf': {'naf_mode': True, 'mel_log': 'none'},
}
MODEL_VARIANTS = {   # tools/gen_golden.py MODEL_VARIANTS: options of model.py:58-72,84-85 on configs/tiny.json
    'elu': {'conv_activation': 'ELU'},
    'act_first': {'relu_after_bn': False},
    'elu_act_first': {'conv_activation': 'ELU', 'relu_after_bn': False},
    'strides': {'strides': [[[1, 2], [2, 1]], [[1, 2], [2, 1]], [[1, 2], [2, 1]], [[1, 2], [2, 1]],
                            [[1, 1], [2, 1]], [[1, 2], [2, 1]], [[1, 1], [2, 1]], [[1, 1], [2, 1]]]},
}


@pytest.mark.parametrize('name', sorted(MEL_VARIANTS))
def test_melspec_variants_vs_reference_golden(golden_dir, name):
    """naf_mode (magnitude, zero padding, slaney scale + norm, + 0.06), log10 / no log, spec_norm = max."""
    g = np.load(os.path.join(golden_dir, 'mel_variants.npz'))[name]
    params = dict(synth.read_config('default'), **MEL_VARIANTS[name])
    y = orc.melspec(_mel_inputs(), params)
    assert y.shape == g.shape == (4, 256, 32)
    err = np.abs(y - g)
    # the all-zero row included: x / max(|x|, 1e-12) = 0 -> the additive constant alone
    tol = 2e-3 if not params.get("naf_mode") else 1e-4      # + 0.06 keeps naf_mode far from the cancellation floor
    assert err.max() < tol, (name, err.max())
    assert err.mean() < 2e-5, (name, err.mean())


@pytest.mark.parametrize('name', sorted(MODEL_VARIANTS))
def test_encoder_variants_vs_reference_golden(golden_dir, name):
    base = synth.read_config('tiny')
    params = dict(base, model=dict(base['model'], **MODEL_VARIANTS[name]))
    g = np.load(os.path.join(golden_dir, 'enc_variants.npz'))
    mel = np.load(os.path.join(golden_dir, 'mel_default.npz'))['mel']
    sd = synth.make_state_dict(params, seed=21)
    z, layers = orc.fpnetwork_forward(sd, mel, params, norm=True, return_layers=True)
    np.testing.assert_allclose(layers[0][1], g['l0_' + name], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(layers[-1][1].reshape(4, -1), g['enc_' + name], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(z, g['z_' + name], rtol=0, atol=2e-5)


def test_frame_pcm16_vs_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'musicdata.npz'))
    for fsm in (1, 2):
        for i, n in enumerate(g['lens']):
            rows = orc.frame_pcm16(synth.synth_pcm(500 + i, int(n)), 8000, 4000 // fsm)
            assert rows.shape[0] == int(g['nseg_f%d_c%d' % (fsm, i)])
            if 'rows_f%d_c%d' % (fsm, i) in g:
                np.testing.assert_allclose(rows, g['rows_f%d_c%d' % (fsm, i)], rtol=0, atol=2e-7)
            else:
                np.testing.assert_allclose(rows[0], g['row0_f%d_c%d' % (fsm, i)], rtol=0, atol=2e-7)
                np.testing.assert_allclose(np.linalg.norm(rows.astype(np.float64), axis=1),
                                           g['rownorm_f%d_c%d' % (fsm, i)], rtol=1e-6)


def test_rerank_vs_reference_golden(golden_dir):
    """database.py:117-166 run unmodified (golden) vs our numpy restatement and vs seq_score."""
    g = np.load(os.path.join(golden_dir, 'db_small.npz'))
    db, key = g['db'], g['key']
    pos = synth.song_pos_from_key(key)
    for c in range(int(g['n_cases'])):
        q, labels, fsm = g['q%d' % c], g['labels%d' % c], int(g['fsm%d' % c])
        sco, (sid, tim), ss = orc.query_embeddings_base(db, pos, q, labels, fsm, 0.5)
        assert sid == int(g['song%d' % c]) and tim == float(g['time%d' % c])
        assert abs(sco - float(g['score%d' % c])) < 1e-6
        np.testing.assert_allclose(ss, g['ss%d' % c], rtol=0, atol=1e-6)
        # the C restatement of cpp/seqscore.cpp gives the same answer through database.py:190-195
        for use_ref in (False, True):
            if use_ref and orc.ref_lib() is None:
                continue
            sco2, (sid2, tim2), ss2 = orc.query_embeddings_cpp(db, pos, q, labels, fsm, 0.5, use_ref=use_ref)
            assert sid2 == sid and tim2 == tim
            assert abs(sco2 - sco) < 1e-6
            np.testing.assert_allclose(ss2, ss, rtol=0, atol=1e-6)


def test_seq_score_restatement_bit_exact_vs_reference_build():
    """Our C restatement == the reference's own seqscore.cpp (oracle/_ref), bit for bit."""
    if orc.ref_lib() is None:
        pytest.skip('oracle/_ref/seqscore.so not built (no /root/reference at build time)')
    db, key = synth.synth_db(5000, d=32, seed=3, song_len=37)
    pos = synth.song_pos_from_key(key)
    qs, songs, offs = synth.synth_queries(db, key, 12, q_len=19, noise=1.0, seed=5)
    for fsm in (1, 2, 3):
        for alpha in (0.0, 4.0):
            for i in range(qs.shape[0]):
                q = qs[i]
                _, labels = orc.flat_ip_search(db, q, 10)
                if i == 3:
                    labels[:, 5:] = -1                       # ragged: fewer than k hits
                a, ssa = orc.seq_score(db, pos, q, labels, fsm, alpha)
                b, ssb = orc.seq_score(db, pos, q, labels, fsm, alpha, use_ref=True)
                assert a == b
                assert np.array_equal(ssa.view(np.uint32), ssb.view(np.uint32))
                if fsm == 1 and alpha == 0.0:
                    assert a == songs[i] and ssa[a, 1] == offs[i]


def test_seq_score_empty_labels():
    db, key = synth.synth_db(100, d=8, seed=1, song_len=10)
    pos = synth.song_pos_from_key(key)
    q = db[:3].copy()
    labels = np.full((3, 4), -1, np.int64)
    best, ss = orc.seq_score(db, pos, q, labels)
    assert best == -1 and not ss.any()
    if orc.ref_lib() is not None:
        best, ss = orc.seq_score(db, pos, q, labels, use_ref=True)
        assert best == -1 and not ss.any()


def test_flat_ip_search_contract():
    db, _ = synth.synth_db(3000, d=24, seed=9)
    q = db[[5, 77, 1234]] + 0.01
    D, I = orc.flat_ip_search(db, q, 7)
    s = (q.astype(np.float64) @ db.T.astype(np.float64))
    ref = np.argsort(-s, axis=1, kind='stable')[:, :7]
    assert np.array_equal(I, ref)
    np.testing.assert_allclose(D, np.take_along_axis(s, ref, 1), rtol=0, atol=1e-6)
    assert (np.diff(D, axis=1) <= 0).all()
    # fewer rows than k -> -1 padded
    D, I = orc.flat_ip_search(db[:4], q, 7)
    assert (I[:, 4:] == -1).all() and (I[:, :4] >= 0).all()
    # exact ties -> lower id first
    dup = np.concatenate([db[:10], db[:10]])
    D, I = orc.flat_ip_search(dup, db[:1], 4)
    assert list(I[0, :2]) == [0, 10]


def test_similarity_loss_vs_reference_golden(golden_dir):
    """oracle.similarity_loss (float64 closed form + analytic gradient) against the reference's own function and
    torch autograd (tools/gen_golden.py runs train.py:41-52 as written)."""
    g = np.load(os.path.join(golden_dir, 'train.npz'))
    for tag, (N, d) in {'n640d64': (640, 64), 'n8d16': (8, 16)}.items():
        rng = np.random.Generator(np.random.PCG64(int(g['seed_' + tag])))
        y = rng.standard_normal((N, d)).astype(np.float32)
        y[1::2] = y[0::2] + 0.5 * y[1::2]
        y /= np.linalg.norm(y, axis=1, keepdims=True)
        loss, dy = orc.similarity_loss(y, 0.05)
        assert abs(loss - float(g['loss_%s_f64' % tag])) < 1e-9
        np.testing.assert_allclose(dy, g['dy_%s_f64' % tag], rtol=0, atol=1e-7)
        assert abs(loss - float(g['loss_' + tag])) < 2e-5                     # the reference in fp32
        np.testing.assert_allclose(dy, g['dy_' + tag], rtol=0, atol=2e-5)


def test_resample_restatement_properties():
    """julius is absent (parity unpinned): the restatement is checked through properties -- identity at equal rates,
    output length, a band-limited sinusoid keeps frequency and amplitude, DC gain 1."""
    t = np.arange(44100) / 44100.0
    x = 0.5 * np.sin(2 * np.pi * 440.0 * t)
    y = orc.resample_frac(x, 44100, 8000)
    assert y.shape[0] == 8000
    want = 0.5 * np.sin(2 * np.pi * 440.0 * np.arange(8000) / 8000.0)
    assert np.abs(y[200:-200] - want[200:-200]).max() < 2e-3
    assert np.array_equal(orc.resample_frac(x, 8000, 8000), x)
    np.testing.assert_allclose(orc.resample_frac(np.ones(4410), 44100, 8000), 1.0, atol=1e-12)
    assert orc.resample_frac(np.zeros((2, 12345)), 16000, 8000).shape == (2, 6172)


def test_mix_mono_fake_stereo_rule():
    rng = np.random.Generator(np.random.PCG64(3))
    a = rng.standard_normal(1000).astype(np.float32)
    assert np.array_equal(orc.mix_mono(np.stack([a, -a])), a)                 # opposite phase: channel 1 is flipped
    b = rng.standard_normal(1000).astype(np.float32)
    assert np.array_equal(orc.mix_mono(np.stack([a, b])), ((a + b) / 2).astype(np.float32))


def test_impulse_response_convolution_vs_reference_fft(golden_dir):
    """oracle.apply_ir (float64 direct form) against the reference's rfft * H_air * H_mic -> irfft -> crop
    (dataset_v2.py:157-163 executed by tools/gen_golden.py): the FFT is long enough that nothing wraps."""
    g = np.load(os.path.join(golden_dir, 'train.npz'))
    x = synth.synth_segments(3, seed=8, seg=8100)
    got = orc.apply_ir(x, [g['ir_air'], g['ir_mic']], 100, 8100)
    assert got.shape == g['ir_out'].shape == (3, 8000)
    assert np.abs(got - g['ir_out']).max() <= 1e-6 * np.abs(g['ir_out']).max()
    # one response, no crop: plain causal convolution truncated to the input length
    one = orc.apply_ir(x[:1, :50], [g['ir_mic'][:1, :7]])
    np.testing.assert_allclose(one[0], np.convolve(x[0, :50].astype(np.float64), g['ir_mic'][0, :7])[:50], atol=1e-12)


def test_mel_kernel_fft_dataflows_reproduce_numpy_rfft():
    """tools/fft_dataflow_check.py emulates, lane by lane, the index algebra of the two mel kernels (mel.cu): the
    five-stage shuffle FFT of the general kernel and the transpose + second in-lane FFT + one butterfly of the fast
    kernel, both with the real-FFT recombination.  Checked against numpy's rfft on the CPU, so a wrong twiddle or
    bit-reversal shows up without a GPU."""
    import importlib.util
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools', 'fft_dataflow_check.py')
    spec = importlib.util.spec_from_file_location('fft_dataflow_check', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(7)
    for _ in range(3):
        x = rng.standard_normal(1024)
        ref = np.fft.rfft(x)
        assert np.abs(ref - mod.warp_fft_1024_real(x)).max() < 1e-10
        assert np.abs(np.abs(ref) ** 2 - mod.warp_fft_1024_real_v2(x)).max() < 1e-12 * (np.abs(ref) ** 2).max()
    assert all(len(b) == 32 for b in mod.transpose_read_banks())      # the transpose reads are bank-conflict free


@pytest.mark.parametrize('act', ['ReLU', 'ELU'])
@pytest.mark.parametrize('act_first', [False, True])
def test_layernorm_activation_backward_vs_torch_autograd(act, act_first):
    """The closed-form backward the training kernels implement (oracle.ln_act_backward: activation derivative from the
    stored output, LayerNorm backward from two per-sample means) against torch autograd on the reference's own layer
    sequence (model.py:58-72), all four option sets."""
    torch = pytest.importorskip('torch')
    rng = np.random.default_rng(11)
    Y = rng.standard_normal((3, 4, 6, 5))
    gamma, beta = 1 + 0.2 * rng.standard_normal((4, 6, 5)), 0.1 * rng.standard_normal((4, 6, 5))
    dA = rng.standard_normal(Y.shape)
    yt = torch.tensor(Y, requires_grad=True)
    gt, bt = torch.tensor(gamma, requires_grad=True), torch.tensor(beta, requires_grad=True)
    f = torch.nn.functional
    a = f.elu if act == 'ELU' else f.relu
    out = f.layer_norm(a(yt), gamma.shape, gt, bt, 1e-5) if act_first else a(f.layer_norm(yt, gamma.shape, gt, bt, 1e-5))
    out.backward(torch.tensor(dA))
    dY, dg, db = orc.ln_act_backward(dA, Y, gamma, beta, act, act_first)
    np.testing.assert_allclose(dY, yt.grad.numpy(), rtol=0, atol=1e-10)
    np.testing.assert_allclose(dg, gt.grad.numpy(), rtol=0, atol=1e-10)
    np.testing.assert_allclose(db, bt.grad.numpy(), rtol=0, atol=1e-10)
