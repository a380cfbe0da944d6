"""GPU (-m gpu): the CUDA path, called through the C-ABI via the reference-shaped Python mirrors, against
the CPU oracle (oracle/) and the golden vectors produced by the reference's own code (tests/golden/)."""
import os

import numpy as np
import pytest

torch = pytest.importorskip('torch')
from oracle import pfann_oracle as orc  # noqa: E402  (checker only)
from pfann_b200 import _lib, synth  # noqa: E402

pytestmark = pytest.mark.gpu

MEL_MAX_TOL = 3e-3     # log domain, bins near the 1e-8 floor: the reference itself (fp32) differs from the double oracle by 1.4e-3 there
MEL_MEAN_TOL = 3e-5
MEL_TOL_ABOVE_FLOOR = 1e-4   # max-abs in the log domain away from the floor (power > 1e-6 of the frame maximum)


def _mel_check(y, ref):
    """SURVEY 8c tolerance: 1e-4 in the log domain for every mel bin whose power exceeds 1e-6 of its frame's maximum
    (cancellation in fp32 only matters below that), 3e-3 for the rest, 3e-5 on average."""
    err = np.abs(y - ref)
    strong = ref > ref.max(axis=-2, keepdims=True) + np.log(1e-6)
    assert strong.mean() > 0.5
    assert err[strong].max() < MEL_TOL_ABOVE_FLOOR, err[strong].max()
    assert err.max() < MEL_MAX_TOL, err.max()
    assert err.mean() < MEL_MEAN_TOL, err.mean()
EMB_FP32_TOL = 1e-4    # abs, unit-norm embeddings, fp32 CUDA-core path
EMB_BF16_COS = 1e-3    # 1 - cos, bf16 tensor-core path (bf16 operands, fp32 accumulate + LayerNorm)


@pytest.fixture(scope='module')
def dev():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda', 0)


def _mel_inputs():
    return np.concatenate([synth.synth_segments(3, seed=1), np.zeros((1, 8000), np.float32)])


# ------------------------------------------------------------------------------------------- stage 1
def test_mel_vs_golden_and_oracle(dev, golden_dir):
    from pfann_b200.datautil.melspec import build_mel_spec_layer
    params = synth.read_config('default')
    mel = build_mel_spec_layer(params).to(dev)
    x = _mel_inputs()
    g = np.load(os.path.join(golden_dir, 'mel_default.npz'))['mel']
    y = mel(torch.from_numpy(x).to(dev)).cpu().numpy()
    assert y.shape == (4, 256, 32)
    for ref in (g, orc.melspec(x, params)):
        err = np.abs(y[:3] - ref[:3])
        assert err.max() < MEL_MAX_TOL, err.max()
        assert err.mean() < MEL_MEAN_TOL, err.mean()
    _mel_check(y[:3], orc.melspec(x, params)[:3])
    np.testing.assert_allclose(y[3], np.log(np.float32(1e-8)), atol=1e-6)      # all-zero segment
    # host pointers go through the same kernel (staged inside the C-ABI call) and give identical bits
    y2 = mel(torch.from_numpy(x)).numpy()
    assert np.array_equal(y, y2)
    # batched / leading dims like the reference module ([..., n] -> [..., n_mels, T])
    y3 = mel(torch.from_numpy(x).to(dev).reshape(2, 2, 8000))
    assert tuple(y3.shape) == (2, 2, 256, 32)
    assert np.array_equal(y3.reshape(4, 256, 32).cpu().numpy(), y)


def test_mel_many_segments_vs_oracle(dev):
    from pfann_b200.datautil.melspec import build_mel_spec_layer
    params = synth.read_config('default')
    mel = build_mel_spec_layer(params).to(dev)
    x = synth.synth_segments(300, seed=3)
    y = mel(torch.from_numpy(x).to(dev)).cpu().numpy()
    _mel_check(y, orc.melspec(x, params))


def test_mel_fast_kernel_equals_the_general_kernel(dev, monkeypatch):
    """The default option set runs the round-2 kernel (transpose FFT, paired recombination, 7-tap projection from
    shared memory); PFANN_B200_MEL_V1 selects the general kernel the variants use.  Same numbers within rounding,
    same moments, other segment lengths included."""
    from pfann_b200.datautil.melspec import build_mel_spec_layer
    params = synth.read_config('default')
    mel = build_mel_spec_layer(params).to(dev)
    for n, count in ((8000, 64), (4000, 5), (12000, 3)):
        x = torch.from_numpy(synth.synth_segments(count, seed=5, seg=n)).to(dev)
        x[-1] = 0
        monkeypatch.delenv('PFANN_B200_MEL_V1', raising=False)
        fast = mel(x).cpu().numpy()
        monkeypatch.setenv('PFANN_B200_MEL_V1', '1')
        slow = mel(x).cpu().numpy()
        monkeypatch.delenv('PFANN_B200_MEL_V1', raising=False)
        ref = orc.melspec(x.cpu().numpy(), params)
        _mel_check(fast[:-1], ref[:-1])
        assert np.abs(fast - slow).mean() < 1e-5
        np.testing.assert_allclose(fast[-1], np.log(np.float32(1e-8)), atol=1e-6)
        assert np.array_equal(fast, mel(x).cpu().numpy())                      # bit-reproducible


MEL_VARIANTS = {     # tools/gen_golden.py MEL_VARIANTS: option sets of melspec.py:27-49 beyond the default
    'naf': {'naf_mode': True, 'mel_log': 'log10', 'spec_norm': 'max'},
    'log10': {'mel_log': 'log10'},
    'max': {'spec_norm': 'max'},
    'nolog_naf': {'naf_mode': True, 'mel_log': 'none'},
}


@pytest.mark.parametrize('name', sorted(MEL_VARIANTS))
def test_mel_variants_vs_golden_and_oracle(dev, golden_dir, name):
    """naf_mode (magnitude, zero padding, slaney scale + norm, + 0.06), log10 / no log, spec_norm = max: the kernel
    against the reference's own output (golden) and the double-precision oracle."""
    from pfann_b200.datautil.melspec import build_mel_spec_layer
    params = dict(synth.read_config('default'), **MEL_VARIANTS[name])
    mel = build_mel_spec_layer(params).to(dev)
    x = _mel_inputs()
    y = mel(torch.from_numpy(x).to(dev)).cpu().numpy()
    g = np.load(os.path.join(golden_dir, 'mel_variants.npz'))[name]
    ref = orc.melspec(x, params)
    naf = bool(params.get('naf_mode'))
    for want in (g, ref):
        err = np.abs(y - want)
        assert err.max() < (2e-4 if naf else MEL_MAX_TOL), (name, err.max())     # + 0.06 keeps naf_mode off the floor
        assert err.mean() < MEL_MEAN_TOL, (name, err.mean())
    # and through the fused PCM entry point (the matcher / builder path): same bits as rows -> mel
    pcm = synth.synth_pcm(77, 12000)
    rows = orc.frame_pcm16(pcm, 8000, 4000)
    h = mel.plan_handle(0, 8000)
    out = np.empty((rows.shape[0], 256, 32), np.float32)
    st, va = np.array([0, 4000], np.int64), np.array([8000, 8000], np.int32)
    _lib.use_torch_stream(0)
    _lib.check(_lib.lib().pfann_mel_forward_pcm16(h, _lib.ptr(pcm), pcm.shape[0], _lib.ptr(st), _lib.ptr(va), 2,
                                                  _lib.ptr(out)))
    err = np.abs(out - orc.melspec(rows, params))
    assert err.max() < (2e-4 if naf else MEL_MAX_TOL) and err.mean() < MEL_MEAN_TOL, (name, err.max(), err.mean())


def test_mel_plans_of_different_segment_lengths_coexist(dev):
    """The layer takes any segment length like the reference module; plans of different lengths (different
    shared-memory sizes) live side by side: long, short, long again."""
    from pfann_b200.datautil.melspec import build_mel_spec_layer
    params = synth.read_config('default')
    mel = build_mel_spec_layer(params).to(dev)
    x = synth.synth_segments(3, seed=8)
    for n in (8000, 4096, 12000, 8000):
        xi = np.ascontiguousarray(np.tile(x, (1, 2))[:, :n])
        y = mel(torch.from_numpy(xi).to(dev)).cpu().numpy()
        ref = orc.melspec(xi, params)
        assert y.shape == ref.shape == (3, 256, 1 + n // 256)
        err = np.abs(y - ref)
        assert err.max() < MEL_MAX_TOL and err.mean() < MEL_MEAN_TOL, (n, err.max(), err.mean())


def test_mel_pcm16_framing_vs_oracle(dev):
    """musicdata.py:48,82-88 folded into the kernel: ragged clips incl. one shorter than a segment."""
    import ctypes
    from pfann_b200.datautil.melspec import build_mel_spec_layer
    params = synth.read_config('default')
    mel = build_mel_spec_layer(params).to(dev)
    lens = [8000, 20000, 5000, 31999]
    pcm = np.concatenate([synth.synth_pcm(500 + i, n) for i, n in enumerate(lens)])
    off = np.concatenate([[0], np.cumsum(lens)])
    for hop in (4000, 2000):
        starts, valids, rows = [], [], []
        for i, n in enumerate(lens):
            r = orc.frame_pcm16(pcm[off[i]:off[i + 1]], 8000, hop)
            rows.append(r)
            for s in range(r.shape[0]):
                starts.append(off[i] + s * hop)
                valids.append(min(8000, n - s * hop))
        rows = np.concatenate(rows)
        B = rows.shape[0]
        out = np.empty((B, 256, 32), np.float32)
        h = mel.plan_handle(0, 8000)
        _lib.use_torch_stream(0)
        st, va = np.array(starts, np.int64), np.array(valids, np.int32)
        _lib.check(_lib.lib().pfann_mel_forward_pcm16(h, _lib.ptr(pcm), pcm.shape[0], _lib.ptr(st), _lib.ptr(va), B,
                                                      _lib.ptr(out)))
        ref = orc.melspec(rows, params)
        err = np.abs(out - ref)
        assert err.max() < MEL_MAX_TOL and err.mean() < MEL_MEAN_TOL, (hop, err.max(), err.mean())


# ------------------------------------------------------------------------------------------- stage 2
def _net(name, precision, dev, seed):
    from pfann_b200.model import FpNetwork
    params = synth.read_config(name)
    d, h, u, F, T = synth.model_dims(params)
    mp = dict(params['model'], b200_precision=precision)
    net = FpNetwork(d, h, u, F, T, mp).to(dev)
    sd = synth.make_state_dict(params, seed=seed)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    net.eval()
    return net, params, sd


@pytest.mark.parametrize('name', ['tiny', 'n640d64', 'default'])
def test_encoder_fp32_vs_golden(dev, golden_dir, name):
    g = np.load(os.path.join(golden_dir, 'enc_%s.npz' % name))
    mel = np.load(os.path.join(golden_dir, 'mel_default.npz'))['mel']
    net, params, sd = _net(name, 'fp32', dev, int(g['seed']))
    x = torch.from_numpy(mel).to(dev)
    z = net(x).cpu().numpy()
    zr = net(x, norm=False).cpu().numpy()
    np.testing.assert_allclose(z, g['z'], rtol=0, atol=EMB_FP32_TOL)
    np.testing.assert_allclose(zr, g['z_raw'], rtol=2e-3, atol=2e-4)
    np.testing.assert_allclose(np.linalg.norm(z, axis=1), 1.0, atol=1e-5)
    # per-layer parity taps against the reference's module outputs (mean / std / max of each layer)
    for l in (0, 3, 7):
        act = net.layer_output(x, l).numpy()
        want = g['layer_stats'][l]
        assert abs(act.mean() - want[0]) < 2e-4 * max(1, abs(want[0])), (l, act.mean(), want)
        assert abs(act.std(ddof=1) - want[1]) < 2e-4 * max(1, want[1]), (l, act.std(ddof=1), want)
    enc = net.layer_output(x, 7).numpy().reshape(x.shape[0], -1)
    np.testing.assert_allclose(enc, g['enc_out'], rtol=2e-3, atol=2e-4)


@pytest.mark.parametrize('name', ['n640d64', 'default'])
def test_encoder_bf16_tensor_core_vs_golden(dev, golden_dir, name):
    g = np.load(os.path.join(golden_dir, 'enc_%s.npz' % name))
    mel = np.load(os.path.join(golden_dir, 'mel_default.npz'))['mel']
    net, params, sd = _net(name, 'bf16', dev, int(g['seed']))
    x = torch.from_numpy(mel).to(dev)
    z = net(x).cpu().numpy()
    cos = (z * g['z']).sum(1)
    assert (1 - cos).max() < EMB_BF16_COS, 1 - cos
    np.testing.assert_allclose(np.linalg.norm(z, axis=1), 1.0, atol=1e-5)
    # layer-by-layer drift stays small (relative L2 error of each SeparableConv2d output vs the fp32 path)
    ref, _, _ = _net(name, 'fp32', dev, int(g['seed']))
    for l in range(8):
        a = net.layer_output(x, l).numpy()
        b = ref.layer_output(x, l).numpy()
        rel = np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-12)
        assert rel < 3e-2, (l, rel)


MODEL_VARIANTS = {   # tools/gen_golden.py MODEL_VARIANTS: options of model.py:58-72,84-85 on configs/tiny.json
    'elu': {'conv_activation': 'ELU'},
    'act_first': {'relu_after_bn': False},
    'elu_act_first': {'conv_activation': 'ELU', 'relu_after_bn': False},
    'strides': {'strides': [[[1, 2], [2, 1]], [[1, 2], [2, 1]], [[1, 2], [2, 1]], [[1, 2], [2, 1]],
                            [[1, 1], [2, 1]], [[1, 2], [2, 1]], [[1, 1], [2, 1]], [[1, 1], [2, 1]]]},
}


@pytest.mark.parametrize('name', sorted(MODEL_VARIANTS))
@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_encoder_variants_vs_golden_and_oracle(dev, golden_dir, name, precision):
    """ELU, activation before the LayerNorm, NAF-style stride schedules: against the reference's own outputs
    (goldens made by model.py) and the oracle.  These options run on the CUDA-core kernels whatever precision is
    asked for, so both settings must give the fp32-grade answer."""
    from pfann_b200.model import FpNetwork
    base = synth.read_config('tiny')
    params = dict(base, model=dict(base['model'], **MODEL_VARIANTS[name]))
    d, h, u, F, T = synth.model_dims(params)
    sd = synth.make_state_dict(params, seed=21)
    net = FpNetwork(d, h, u, F, T, dict(params['model'], b200_precision=precision)).to(dev)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    net.eval()
    g = np.load(os.path.join(golden_dir, 'enc_variants.npz'))
    mel = np.load(os.path.join(golden_dir, 'mel_default.npz'))['mel']
    x = torch.from_numpy(mel).to(dev)
    z = net(x).cpu().numpy()
    np.testing.assert_allclose(z, g['z_' + name], rtol=0, atol=EMB_FP32_TOL)
    np.testing.assert_allclose(net.layer_output(x, 0).numpy(), g['l0_' + name], rtol=2e-3, atol=2e-4)
    np.testing.assert_allclose(net.layer_output(x, 7).numpy().reshape(4, -1), g['enc_' + name], rtol=2e-3, atol=2e-4)
    np.testing.assert_allclose(z, orc.fpnetwork_forward(sd, mel, params), rtol=0, atol=EMB_FP32_TOL)
    # default-config extract path with a variant model: PCM in, fingerprints out
    zr = net(x, norm=False).cpu().numpy()
    np.testing.assert_allclose(zr / np.linalg.norm(zr, axis=1, keepdims=True), z, rtol=0, atol=1e-5)


def test_encoder_bf16_is_deterministic_and_chunk_invariant(dev):
    """LayerNorm partial sums use fixed slots (no atomics): identical bits run to run, and the result does
    not depend on how segments are grouped into chunks."""
    net, params, sd = _net('default', 'bf16', dev, 5)
    x = torch.from_numpy(orc.melspec(synth.synth_segments(37, seed=9), params)).to(dev)
    z1 = net(x).cpu().numpy()
    z2 = net(x).cpu().numpy()
    assert np.array_equal(z1, z2)
    net.chunk = 16
    z3 = net(x).cpu().numpy()
    assert np.array_equal(z1, z3)
    ref = orc.fpnetwork_forward(sd, x[:2].cpu().numpy(), params)
    assert (1 - (z1[:2] * ref).sum(1)).max() < EMB_BF16_COS


def test_layer0_fusion_matches_unfused(dev, monkeypatch):
    """conv1 + ln1 + ReLU of layer 0 with statistics derived from the mel moments == conv -> two-pass LN."""
    params = synth.read_config('default')
    x = torch.from_numpy(orc.melspec(synth.synth_segments(5, seed=2), params)).to(dev)
    fused, _, _ = _net('default', 'fp32', dev, 7)
    a = fused.layer_output(x, 0).numpy()
    za = fused(x).cpu().numpy()
    monkeypatch.setenv('PFANN_B200_NO_L0_FUSION', '1')
    plain, _, _ = _net('default', 'fp32', dev, 7)
    b = plain.layer_output(x, 0).numpy()
    zb = plain(x).cpu().numpy()
    np.testing.assert_allclose(a, b, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(za, zb, rtol=0, atol=2e-5)


def test_fused_conv_layernorm_matches_unfused(dev, monkeypatch):
    """conv_ln_tc_kernel (LayerNorm inside the GEMM epilogue, statistics exchanged between CTAs) against the
    conv -> partial sums -> finalize -> apply chain on the same bf16 operands, for a batch that spans several
    tiles per CTA position and is not a multiple of the samples-per-tile of the late layers."""
    net, params, sd = _net('default', 'bf16', dev, 5)
    x = torch.from_numpy(orc.melspec(synth.synth_segments(43, seed=12), params)).to(dev)
    fused = [net.layer_output(x, l).numpy() for l in range(8)]
    zf = net(x).cpu().numpy()
    monkeypatch.setenv('PFANN_B200_NO_FUSED_LN', '1')
    plain = [net.layer_output(x, l).numpy() for l in range(8)]
    zp = net(x).cpu().numpy()
    for l in range(8):
        rel = np.linalg.norm(fused[l] - plain[l]) / np.linalg.norm(plain[l])
        assert rel < 2e-2, (l, rel)      # both round activations to bf16, at different points of the arithmetic
    assert (1 - (zf * zp).sum(1)).max() < 2e-4
    monkeypatch.delenv('PFANN_B200_NO_FUSED_LN')
    assert np.array_equal(net(x).cpu().numpy(), zf)     # and the fused path is bit-reproducible


def test_front_kernel_matches_two_kernel_layer0(dev, monkeypatch):
    """front_tc_kernel (layer-0 conv1 + ln1 + ReLU produced inside conv2's operand pipeline, X0 never stored) against
    the round-1 pair l0_tc_kernel -> conv_ln_tc_kernel that materialises X0: same layer-0 output up to bf16 rounding
    of X0, for a batch that is not a multiple of the 8 segments of a tile, over several chunks, bit-reproducible."""
    net, params, sd = _net('default', 'bf16', dev, 5)
    x = torch.from_numpy(orc.melspec(synth.synth_segments(43, seed=21), params)).to(dev)
    a0 = net.layer_output(x, 0).numpy()
    za = net(x).cpu().numpy()
    net.chunk = 16
    assert np.array_equal(net(x).cpu().numpy(), za)               # independent of the grouping into tiles / chunks
    assert np.array_equal(net.layer_output(x[:16], 0).numpy(), a0[:16])
    monkeypatch.setenv('PFANN_B200_NO_FRONT', '1')
    b0 = net.layer_output(x[:16], 0).numpy()
    zb = net(x).cpu().numpy()
    rel = np.linalg.norm(a0[:16] - b0) / np.linalg.norm(b0)
    assert rel < 5e-3, rel
    assert (1 - (za * zb).sum(1)).max() < 1e-4
    monkeypatch.delenv('PFANN_B200_NO_FRONT')
    ref = orc.fpnetwork_forward(sd, x[:3].cpu().numpy(), params)
    assert (1 - (za[:3] * ref).sum(1)).max() < EMB_BF16_COS


def test_layer0_tensor_core_matches_cuda_core(dev, monkeypatch):
    """l0_tc_kernel (conv1 of layer 0 as one K = 16 bf16 hi/lo split MMA per 128 positions) against the CUDA-core
    fp32 formulation: same statistics, same bf16 output up to rounding."""
    net, params, sd = _net('default', 'bf16', dev, 9)
    x = torch.from_numpy(orc.melspec(synth.synth_segments(6, seed=4), params)).to(dev)
    a = net.layer_output(x, 0).numpy()
    monkeypatch.setenv('PFANN_B200_NO_L0_TC', '1')
    b = net.layer_output(x, 0).numpy()
    assert np.linalg.norm(a - b) / np.linalg.norm(b) < 5e-3


def test_extract_host_buffers_pipeline_equals_device_buffers(dev):
    """pfann_extract_pcm16 with host buffers overlaps H2D / compute / D2H chunk by chunk; the fingerprints must be
    the bits the device-buffer call produces (several chunks, ragged clip lengths, a clip shorter than a segment)."""
    from pfann_b200.extract import Extractor
    params = synth.read_config('default')
    ex = Extractor(params, synth.make_state_dict(params, seed=3), device=0, precision='bf16', chunk=16)
    lens = [52000, 8000, 3000, 91000, 24000, 40001]
    pcm = np.concatenate([synth.synth_pcm(300 + i, n) for i, n in enumerate(lens)])
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    z_host, c_host = ex.extract_pcm16(pcm, off)
    z_dev, c_dev = ex.extract_pcm16(torch.from_numpy(pcm).to(dev), off)
    assert z_host.shape[0] > 3 * 16 and np.array_equal(c_host, c_dev)
    assert np.array_equal(z_host, z_dev.cpu().numpy())
    pinned = torch.from_numpy(pcm).pin_memory()
    out = torch.empty(z_host.shape, dtype=torch.float32).pin_memory()
    ex.extract_pcm16(pinned, off, out=out)
    assert np.array_equal(out.numpy(), z_host)


def test_models_of_different_configs_and_precisions_interleave(dev):
    """Regression: launch-attribute caches are per process -- a small fp32 model used after a large bf16 model once
    lowered the head kernel's shared-memory opt-in and the next large launch failed with 'invalid argument'."""
    big, params, _ = _net('default', 'bf16', dev, 3)
    x = torch.from_numpy(orc.melspec(synth.synth_segments(3, seed=1), params)).to(dev)
    z0 = big(x).cpu().numpy()
    small32, _, _ = _net('n640d64', 'fp32', dev, 3)
    small16, _, _ = _net('n640d64', 'bf16', dev, 3)
    assert small32(x).shape == (3, 64) and small16(x).shape == (3, 64)
    big32, _, _ = _net('default', 'fp32', dev, 3)
    z1 = big(x).cpu().numpy()
    z2 = big32(x).cpu().numpy()
    assert np.array_equal(z0, z1)
    assert (1 - (z1 * z2).sum(1)).max() < EMB_BF16_COS


def test_extract_pcm16_equals_mel_plus_model(dev):
    """builder.py:88-99 fused: PCM in, fingerprints out == framing -> mel -> model done step by step."""
    import ctypes
    from pfann_b200.datautil.melspec import build_mel_spec_layer
    net, params, sd = _net('n640d64', 'bf16', dev, 21)
    mel = build_mel_spec_layer(params).to(dev)
    lens = [30000, 8000, 5000, 44000]
    pcm = np.concatenate([synth.synth_pcm(900 + i, n) for i, n in enumerate(lens)])
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    L = _lib.lib()
    nseg = L.pfann_count_segments(off.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), len(lens), 8000, 4000)
    z = np.empty((nseg, 64), np.float32)
    counts = np.empty(len(lens), np.int32)
    _lib.use_torch_stream(0)
    _lib.check(L.pfann_extract_pcm16(mel.plan_handle(0, 8000), net.native_handle(0), _lib.ptr(pcm),
                                     off.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), len(lens), 4000, 1,
                                     _lib.ptr(z), counts.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))))
    rows = np.concatenate([orc.frame_pcm16(pcm[off[i]:off[i + 1]], 8000, 4000) for i in range(len(lens))])
    assert list(counts) == [(max(n, 8000) - 8000) // 4000 + 1 for n in lens] and counts.sum() == nseg
    z2 = net(mel(torch.from_numpy(rows).to(dev))).cpu().numpy()
    assert (1 - (z * z2).sum(1)).max() < 1e-4       # same kernels; only the mean-removal rounding differs
    z3 = np.empty_like(z)
    _lib.check(L.pfann_extract_segments(mel.plan_handle(0, 8000), net.native_handle(0), _lib.ptr(rows), nseg, 1,
                                        _lib.ptr(z3)))
    assert np.array_equal(z3, z2)


# ------------------------------------------------------------------------------------------- stage 3
def _open_db(tmp_path, db, key, top_k, fsm=1, alpha=0):
    from pfann_b200.database import Database, write_flat_ip_index
    d = str(tmp_path)
    db.tofile(os.path.join(d, 'embeddings'))
    write_flat_ip_index(os.path.join(d, 'landmarkValue'), db)
    np.asarray(key, np.int32).tofile(os.path.join(d, 'landmarkKey'))
    with open(os.path.join(d, 'songList.txt'), 'w') as f:
        f.write('\n'.join('song%d.wav' % i for i in range(len(key))) + '\n')
    ix = {'top_k': top_k, 'frame_shift_mul': fsm}
    if alpha:
        ix['score_alpha'] = alpha
    return Database(d, ix, 0.5)


def _check_topk(D, I, Dref, Iref, db, q):
    """Exact-search parity: distances bit-exact for equal labels; label differences only at exact ties."""
    assert D.shape == Dref.shape
    same = I == Iref
    assert np.array_equal(D[same].view(np.uint32), Dref[same].view(np.uint32))
    if not same.all():
        # any mismatch must be a tie in score (both orders are then valid rankings up to the id rule)
        assert np.array_equal(D.view(np.uint32), Dref.view(np.uint32))
        assert False, 'tie order differs from (score desc, id asc)'


@pytest.mark.parametrize('use_tc', [0, 1])
@pytest.mark.parametrize('k', [1, 20, 100])
def test_knn_exact_vs_oracle(dev, tmp_path, use_tc, k):
    db, key = synth.synth_db(50000, d=128, seed=4)
    qs, _, _ = synth.synth_queries(db, key, 3, q_len=19, seed=8)
    q = qs.reshape(-1, 128)
    dbo = _open_db(tmp_path, db, key, k)
    _lib.check(_lib.lib().pfann_db_set_tuning(dbo.handle, 0, 0, use_tc))
    D, I = dbo.search(q)
    Dref, Iref = orc.flat_ip_search(db, q, k)
    _check_topk(D, I, Dref, Iref, db, q)
    assert (np.diff(D, axis=1) <= 0).all()


@pytest.mark.parametrize('use_tc', [0, 1])
def test_knn_edge_cases(dev, tmp_path, use_tc):
    # fewer rows than k -> -1 / -FLT_MAX padding (faiss contract, database.py:135)
    db, key = synth.synth_db(7, d=128, seed=1, song_len=7)
    dbo = _open_db(tmp_path, db, key, 10)
    _lib.check(_lib.lib().pfann_db_set_tuning(dbo.handle, 0, 0, use_tc))
    D, I = dbo.search(db[:3])
    Dref, Iref = orc.flat_ip_search(db, db[:3], 10)
    assert np.array_equal(I, Iref) and np.array_equal(D.view(np.uint32), Dref.view(np.uint32))
    assert (I[:, 7:] == -1).all()
    dbo.close()
    # duplicated rows: exact ties resolve to the lower id; tiny candidate lists force the overflow/backstop paths
    base, _ = synth.synth_db(600, d=128, seed=2)
    dup = np.concatenate([base, base, base])
    key = np.full(dup.shape[0] // 100, 100, np.int32)
    dbo = _open_db(tmp_path, dup, key, 8)
    _lib.check(_lib.lib().pfann_db_set_tuning(dbo.handle, 16, 64, use_tc))
    q = base[:40] + 0.01
    D, I = dbo.search(q)
    Dref, Iref = orc.flat_ip_search(dup, q, 8)
    assert np.array_equal(I, Iref) and np.array_equal(D.view(np.uint32), Dref.view(np.uint32))
    dbo.close()
    # all-zero database: every score ties
    z = np.zeros((300, 128), np.float32)
    dbo = _open_db(tmp_path, z, np.full(3, 100, np.int32), 5)
    _lib.check(_lib.lib().pfann_db_set_tuning(dbo.handle, 8, 32, use_tc))
    D, I = dbo.search(base[:2])
    assert np.array_equal(I, np.tile(np.arange(5), (2, 1))) and (D == 0).all()


def test_knn_many_queries_batched(dev, tmp_path):
    """> 128 queries per call (several database passes), d = 64 (n640d64), against the oracle."""
    db, key = synth.synth_db(20000, d=64, seed=6)
    q = synth.synth_queries(db, key, 20, q_len=19, seed=3)[0].reshape(-1, 64)
    dbo = _open_db(tmp_path, db, key, 20)
    D, I = dbo.search(q)
    Dref, Iref = orc.flat_ip_search(db, q, 20)
    _check_topk(D, I, Dref, Iref, db, q)


def test_knn_256_queries_per_pass(dev, tmp_path):
    """d = 128 and 300 queries: one 256-query pass (two TMEM chunks per epilogue warp) + one 44-query pass."""
    db, key = synth.synth_db(60000, d=128, seed=16)
    q = synth.synth_queries(db, key, 16, q_len=19, seed=5)[0].reshape(-1, 128)[:300]
    dbo = _open_db(tmp_path, db, key, 20)
    D, I = dbo.search(q)
    Dref, Iref = orc.flat_ip_search(db, q, 20)
    _check_topk(D, I, Dref, Iref, db, q)


def test_seq_score_bit_exact_vs_reference_build(dev, tmp_path, golden_dir):
    """The GPU rerank behind the reference's own seq_score() signature vs the reference's own seqscore.cpp
    (oracle/_ref) and our C restatement: song id, offsets and scores bit for bit (score_alpha = 0)."""
    use_ref = orc.ref_lib() is not None
    g = np.load(os.path.join(golden_dir, 'db_small.npz'))
    db, key = g['db'], g['key']
    pos = synth.song_pos_from_key(key)
    for c in range(int(g['n_cases'])):
        q, labels, fsm = g['q%d' % c], g['labels%d' % c], int(g['fsm%d' % c])
        dbo = _open_db(tmp_path, db, key, labels.shape[1], fsm=fsm)
        ss = np.zeros((len(key), 2), np.float32)
        import ctypes
        from ctypes import POINTER, c_float, c_int64
        best = _lib.lib().seq_score(dbo.handle, pos.ctypes.data_as(POINTER(c_int64)), len(key),
                                    q.ctypes.data_as(POINTER(c_float)), q.shape[0],
                                    labels.ctypes.data_as(POINTER(c_int64)), labels.shape[1],
                                    ss.ctypes.data_as(POINTER(c_float)), fsm, 0.0)
        rbest, rss = orc.seq_score(db, pos, q, labels, fsm, 0.0, use_ref=use_ref)
        assert best == rbest
        assert np.array_equal(ss.view(np.uint32), rss.view(np.uint32))
        # and the reference's Python path (database.py:117-166, golden) agrees on (song, time)
        sco, (sid, tim), full = dbo.query_embeddings(q)
        assert sid == int(g['song%d' % c]) and tim == float(g['time%d' % c])
        assert abs(sco - float(g['score%d' % c])) < 1e-6
        np.testing.assert_allclose(full, g['ss%d' % c], rtol=0, atol=1e-6)
        dbo.close()


@pytest.mark.parametrize('fsm,alpha', [(1, 0.0), (2, 0.0), (3, 0.0), (1, 4.0)])
def test_database_query_vs_oracle(dev, tmp_path, fsm, alpha):
    db, key = synth.synth_db(30000, d=128, seed=11, song_len=59)
    pos = synth.song_pos_from_key(key)
    qs, songs, offs = synth.synth_queries(db, key, 16, q_len=19 * fsm if fsm > 1 else 19, seed=5)
    dbo = _open_db(tmp_path, db, key, 20, fsm=fsm, alpha=alpha)
    use_ref = orc.ref_lib() is not None
    flat, index = [], []
    for i in range(qs.shape[0]):
        q = qs[i]
        sco, (sid, tim), ss = dbo.query_embeddings(q)
        _, labels = orc.flat_ip_search(db, q, 20)
        rsco, (rsid, rtim), rss = orc.query_embeddings_cpp(db, pos, q, labels, fsm, 0.5, alpha, use_ref=use_ref)
        assert sid == rsid and tim == rtim
        if alpha == 0.0:
            assert sco == rsco and np.array_equal(ss.view(np.uint32), rss.view(np.uint32))
        else:
            assert abs(sco - rsco) < 1e-5
            np.testing.assert_allclose(ss, rss, rtol=0, atol=1e-5)
        if fsm == 1:
            assert sid == songs[i] and tim == offs[i] * 0.5      # planted diagonal recovered
        index.append([len(flat) * q.shape[0], q.shape[0]])
        flat.append(q)
    # batched form == per-file form
    bs, bsong, btime, bss = dbo.query_batch(np.concatenate(flat), np.array(index), want_song_scores=True)
    for i in range(qs.shape[0]):
        sco, (sid, tim), ss = dbo.query_embeddings(qs[i])
        assert bsong[i] == sid and btime[i] == tim
        assert np.float32(sco) == bs[i]
        assert np.array_equal(bss[i].view(np.uint32), ss.view(np.uint32))


def test_query_edge_cases(dev, tmp_path):
    # empty database (database.py:126-127) and queries that overhang song ends / hit zero-length songs
    from pfann_b200.database import Database
    db, key = synth.synth_db(0, d=128, seed=1, song_len=5) if False else (np.zeros((0, 128), np.float32), np.zeros(2, np.int32))
    dbo = _open_db(tmp_path, db, key, 5)
    D, I = dbo.search(np.ones((2, 128), np.float32))
    assert (I == -1).all()
    sco, (sid, tim), ss = dbo.query_embeddings(np.ones((3, 128), np.float32))
    assert sid == -1 and sco == 0.0 and tim == 0.0 and not ss.any()


def test_two_devices_in_one_process(dev):
    """The shared-memory opt-in of every kernel is per device: a second GPU used from the same process must work
    (a process-wide cache of cudaFuncSetAttribute once skipped it there) and give the same fingerprints."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two CUDA devices')
    from pfann_b200.extract import Extractor
    params = synth.read_config('default')
    sd = synth.make_state_dict(params, seed=3)
    lens = [20000, 8000, 12000]
    pcm = np.concatenate([synth.synth_pcm(70 + i, n) for i, n in enumerate(lens)])
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    z0, _ = Extractor(params, sd, device=0, precision='bf16', chunk=16).extract_pcm16(pcm, off)
    z1, _ = Extractor(params, sd, device=1, precision='bf16', chunk=16).extract_pcm16(pcm, off)
    assert np.array_equal(z0, z1)
    db, key = synth.synth_db(5000, d=128, seed=4)
    from pfann_b200.database import Database
    a = Database.from_arrays(db, key, {'top_k': 10, 'frame_shift_mul': 1}, 0.5, device=0)
    b = Database.from_arrays(db, key, {'top_k': 10, 'frame_shift_mul': 1}, 0.5, device=1)
    q = db[100:119]
    assert a.query_embeddings(q)[:2] == b.query_embeddings(q)[:2]


def test_sharded_search_phases_three_shards_one_gpu(dev):
    """The multi-GPU exchange (pfann_b200/dist.py) with its collectives done by hand on one GPU: three shards cut at
    song boundaries (one of them empty), thresholds combined with an element-wise maximum, ONE list of packed keys per
    shard, merge, shard-local sequence score, winner combination == the unsharded database, bit for bit.  The global
    thresholds leave shards with fewer than k survivors, and the deferred overflow check stays at zero."""
    from pfann_b200.database import Database
    from pfann_b200.dist import GpuShard, ShardedDatabase, shard_songs
    db, key = synth.synth_db(60000, d=128, seed=4, song_len=59)
    key = np.concatenate([key, np.zeros(3, np.int32)])             # trailing empty songs -> an empty last shard
    pos = synth.song_pos_from_key(key)
    qs, songs, offs = synth.synth_queries(db, key[:-3], 70, q_len=19, seed=8)
    q_flat = qs.reshape(-1, 128)
    qi = np.stack([np.arange(70) * 19, np.full(70, 19)], 1).astype(np.int64)
    k = 20
    params = {'top_k': k, 'frame_shift_mul': 1}
    full = Database.from_arrays(db, key, params, 0.5, device=0)
    cuts = shard_songs(pos, 2) + [(len(key), len(key))]
    shards = [GpuShard(Database.from_arrays(db, key, params, 0.5, device=0, songs=c)) for c in cuts]
    m = max(s.max_norm() for s in shards)
    for s in shards:
        s.set_max_norm(m)
    q = shards[0].to_device(q_flat)
    thr_max = torch.stack([s.thresholds(q, k) for s in shards]).amax(dim=0)
    assert torch.isinf(shards[2].thresholds(q, k)).all()            # an empty shard bounds nothing
    # thresholds from the union of the shards' samples (the exchange dist.py uses): still a lower bound of the global
    # k-th best score, at least as tight as the maximum of the per-shard bounds, identical on every shard
    top = torch.stack([s.sample_topk(q, k) for s in shards])
    assert (top[2] == 0).all()
    thr = shards[0].thresholds_from_topk(top, q, k)
    assert torch.equal(thr, shards[2].thresholds_from_topk(top, q, k))
    kth = torch.from_numpy(np.sort(q_flat @ db.T, axis=1)[:, ::-1][:, k - 1].copy()).to(thr.device)
    assert (thr <= kth).all() and (thr >= thr_max - 1e-6).all() and (thr > thr_max).any()
    keys = torch.stack([s.filtered_keys(q, k, thr.clone(), True) for s in shards])
    assert (keys[2] == 0).all() and (keys[:2] == 0).any()           # empty shard; some lists shorter than k
    D, I = shards[0].merge_keys(keys, k, want_dist=True)
    D0, I0 = full.search(q_flat, k)
    assert np.array_equal(I.cpu().numpy(), I0) and np.array_equal(D.cpu().numpy().view(np.uint32), D0.view(np.uint32))
    packed = torch.stack([s.rerank_packed(q, qi, I, k, 1, 0.0) for s in shards])
    out = shards[1].combine(packed).cpu().numpy()
    rs, rg, rt, _ = full.query_batch(q_flat, qi)
    assert np.array_equal(out[:, 0], rs) and np.array_equal(np.ascontiguousarray(out[:, 1]).view(np.int32), rg)
    assert np.array_equal(out[:, 2].astype(np.float64) * 0.5, rt)
    assert np.array_equal(rg, songs) and np.array_equal(rt, offs * 0.5)
    torch.cuda.synchronize()
    assert all(s.take_overflow() == 0 for s in shards)
    # the single-process wrapper (world size 1: no collectives) gives the same answers, batched or not
    sdb = ShardedDatabase(GpuShard(full), k, 1, 0.5)
    for res in (sdb.query_batch(q_flat, qi), sdb.query_batches(q_flat, qi, 32)):
        assert np.array_equal(res[0], rs) and np.array_equal(res[1], rg) and np.array_equal(res[2], rt)
    # a tiny candidate capacity forces overflows: counted in the deferred mode, repaired by the checked rerun
    _lib.check(_lib.lib().pfann_db_set_tuning(full.handle, 32, -1, -1))
    res = sdb.query_batches(q_flat, qi, 32)
    assert np.array_equal(res[0], rs) and np.array_equal(res[1], rg) and np.array_equal(res[2], rt)
