"""GPU (-m gpu): SURVEY 8f.2 / 8f.3 pieces -- NT-Xent loss forward + backward, SpecAugment, SNR mix (train.py,
datautil/specaug.py, datautil/noise.py) against goldens made by the reference's own functions and the oracle; GPU
ingest (planar conversion, fractional resampling, mono mix with the fake-stereo rule) against the oracle."""
import os
import wave

import numpy as np
import pytest

torch = pytest.importorskip('torch')
from oracle import pfann_oracle as orc  # noqa: E402  (checker only)
from pfann_b200 import synth  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda', 0)


def _pairs(N, d, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    y = rng.standard_normal((N, d)).astype(np.float32)
    y[1::2] = y[0::2] + 0.5 * y[1::2]
    y /= np.linalg.norm(y, axis=1, keepdims=True)
    return y


@pytest.mark.parametrize('tag,N,d', [('n640d64', 640, 64), ('n8d16', 8, 16)])
def test_similarity_loss_forward_backward_vs_reference(dev, golden_dir, tag, N, d):
    """train.py:41-52: loss and dL/dy to 1e-5 of the reference's own function under torch autograd (golden) and of
    the float64 oracle; bit-reproducible; the gradient flows through torch autograd like the reference's."""
    from pfann_b200.train import similarity_loss
    g = np.load(os.path.join(golden_dir, 'train.npz'))
    y = torch.from_numpy(_pairs(N, d, int(g['seed_' + tag]))).to(dev).requires_grad_(True)
    loss = similarity_loss(y, 0.05)
    (2.0 * loss).backward()
    lo, dy = orc.similarity_loss(y.detach().cpu().numpy(), 0.05)
    assert abs(loss.item() - lo) < 1e-5 and abs(loss.item() - float(g['loss_' + tag])) < 1e-5
    got = y.grad.cpu().numpy() / 2.0
    np.testing.assert_allclose(got, dy, rtol=0, atol=1e-5)
    np.testing.assert_allclose(got, g['dy_' + tag], rtol=0, atol=1e-5)
    y2 = y.detach().clone().requires_grad_(True)
    l2 = similarity_loss(y2, 0.05)
    l2.backward()
    assert l2.item() == loss.item() and torch.equal(y2.grad * 2.0, y.grad)
    # n640 batch with d = 128 (default.json's fingerprint size) and another temperature
    if N == 640:
        y3 = torch.from_numpy(_pairs(640, 128, 9)).to(dev).requires_grad_(True)
        l3 = similarity_loss(y3, 0.1)
        l3.backward()
        lo3, dy3 = orc.similarity_loss(y3.detach().cpu().numpy(), 0.1)
        assert abs(l3.item() - lo3) < 1e-5
        np.testing.assert_allclose(y3.grad.cpu().numpy(), dy3, rtol=0, atol=1e-5)


def test_specaugment_masks_follow_the_reference_rng(dev, golden_dir):
    """Same torch seed -> the masks SpecAugment.get_mask draws (golden: specaug.py run as is) and the same augmented
    batch, applied by the CUDA kernel."""
    from pfann_b200.train import SpecAugment
    g = np.load(os.path.join(golden_dir, 'train.npz'))
    sa = SpecAugment({'cutout_min': 0.1, 'cutout_max': 0.5})
    torch.manual_seed(1234)
    masks = np.stack([sa.get_mask(256, 32).numpy() for _ in range(6)]).astype(np.uint8)
    assert np.array_equal(masks, g['specaug_masks'])
    torch.manual_seed(77)
    x = (torch.arange(2 * 256 * 32, dtype=torch.float32).reshape(2, 256, 32) + 1).to(dev)
    assert np.array_equal(sa.augment(x).cpu().numpy(), g['specaug_x'])
    # per-sample masks: sample b gets the b-th mask of the stream
    torch.manual_seed(1234)
    xb = torch.ones((6, 256, 32), device=dev)
    assert np.array_equal(sa.augment_batch(xb).cpu().numpy(), 1.0 - g['specaug_masks'].astype(np.float32))


def test_snr_mix_vs_oracle(dev):
    from pfann_b200.train import add_noises
    rng = np.random.Generator(np.random.PCG64(4))
    x = rng.standard_normal((7, 9600)).astype(np.float32) * 0.1
    nz = rng.standard_normal((7, 9600)).astype(np.float32) * 0.3
    nz[3] = 0.0                                                    # silent noise row: the clamp decides
    snr = np.linspace(-6, 8, 7).astype(np.float32)
    got = add_noises(torch.from_numpy(x).to(dev), torch.from_numpy(nz), snr).cpu().numpy()
    np.testing.assert_allclose(got, orc.add_noises(x, nz, snr), rtol=1e-5, atol=1e-6)


def _wav(path, pcm, nch, rate):
    with wave.open(path, 'wb') as w:
        w.setnchannels(nch)
        w.setsampwidth(2)
        w.setframerate(rate)
        w.writeframes(np.ascontiguousarray(pcm).tobytes())


def test_gpu_ingest_resample_and_mix_vs_oracle(dev, tmp_path):
    """musicdata.py:29-80 on the GPU: stereo 44.1 kHz / mono 16 kHz / fake-stereo 8 kHz clips -> mono at 8 kHz ==
    the oracle's restatement (resampler: julius algorithm, unpinned); then the fused extract gives the fingerprints
    of the same samples fed through the float-row entry point."""
    from pfann_b200.datautil import musicdata
    from pfann_b200.extract import Extractor
    params = synth.read_config('default')
    ex = Extractor(params, synth.make_state_dict(params, seed=11), device=0, precision='bf16', chunk=64)
    rng = np.random.Generator(np.random.PCG64(12))
    t44 = np.arange(int(2.7 * 44100)) / 44100.0
    left = 0.4 * np.sin(2 * np.pi * 523.0 * t44) + 0.05 * rng.standard_normal(t44.shape)
    right = 0.3 * np.sin(2 * np.pi * 1310.0 * t44 + 1.0) + 0.05 * rng.standard_normal(t44.shape)
    stereo = np.round(np.stack([left, right], 1) * 32767 * 0.5).astype(np.int16)
    t16 = np.arange(16000 * 3) / 16000.0
    mono16 = np.round(0.5 * np.sin(2 * np.pi * 700.0 * t16) * 32767 * 0.5).astype(np.int16)[:, None]
    m8 = synth.synth_pcm(5, 20000)
    fake = np.stack([m8, -m8], 1)                                  # opposite phase: must not cancel
    cases = [(stereo, 44100), (mono16, 16000), (fake, 8000)]
    monos = []
    for pcm, rate in cases:
        got = ex.ingest_wav(pcm, rate).cpu().numpy()
        x = np.multiply(pcm.T, 1 / 32768, dtype=np.float32)
        want = orc.mix_mono(orc.resample_frac(x, rate, 8000).astype(np.float32))
        assert got.shape == want.shape == (int(pcm.shape[0] * 8000 // rate),)
        np.testing.assert_allclose(got, want, rtol=0, atol=2e-6)
        monos.append(got)
    assert np.abs(monos[2] - m8.astype(np.float32) / 32768).max() < 1e-7      # fake stereo recovered, not silence
    # files -> fingerprints through the command-line reader: same as framing the ingested samples on the host
    paths = []
    for i, (pcm, rate) in enumerate(cases):
        p = str(tmp_path / ('c%d.wav' % i))
        _wav(p, pcm, pcm.shape[1], rate)
        paths.append(p)
    wavs = [musicdata.read_wav_pcm16(p, 8000) for p in paths]
    assert [k for k, _ in wavs] == ['wav', 'wav', 'wav']
    z, counts = ex.extract_wavs([d for _, d in wavs])
    assert list(counts) == [4, 5, 4]
    rows = np.concatenate([musicdata.frame_float(m, 8000, 4000) for m in monos])
    zr = ex.extract_segments(rows)
    assert (1 - (z * zr).sum(1)).max() < 1e-5


def _train_net(name, dev):
    from pfann_b200.model import FpNetwork
    cfg, opt, B, seed = synth.TRAIN_CASES[name]
    base = synth.read_config(cfg)
    params = dict(base, model=dict(base['model'], **opt))
    d, h, u, F, T = synth.model_dims(params)
    net = FpNetwork(d, h, u, F, T, params['model']).to(dev)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_state_dict(params, seed=seed).items()})
    return net.train(), params, seed


@pytest.mark.parametrize('name', list(synth.TRAIN_CASES))
def test_training_step_gradients_vs_reference_autograd(dev, golden_dir, name):
    """train.py:96-103: z = model(x); loss = similarity_loss(z, tau); loss.backward().  Every parameter's gradient
    against the reference network under torch autograd (golden: l2 norm + a seeded sample of elements), driven by the
    reference's own dL/dz so that this checks the encoder backward alone.  Cases: dense and depthwise conv2, a custom
    stride schedule, n640d64, and the option variants ELU / activation before the LayerNorm / both (model.py:58-72)."""
    g = np.load(os.path.join(golden_dir, 'train_step.npz'))
    net, params, seed = _train_net(name, dev)
    x = torch.from_numpy(synth.train_case_input(name)).to(dev)
    z = net(x)
    assert z.requires_grad
    np.testing.assert_allclose(z.detach().cpu().numpy(), g[name + '/z'], rtol=0, atol=2e-5)
    z.backward(torch.from_numpy(g[name + '/dz']).to(dev))
    worst = 0.0
    for i, (k, p) in enumerate(net.named_parameters()):
        assert p.grad is not None and p.grad.shape == p.shape, k
        got = p.grad.detach().cpu().numpy().reshape(-1)
        ref_norm = float(g['%s/norm/%s' % (name, k)])
        ref_vals = g['%s/vals/%s' % (name, k)]
        vals = got[synth.grad_sample_index(got.size, seed * 100 + i)]
        # fp32 sums in another order than torch's: 1e-4 of the gradient's scale (its rms) plus 1e-4 relative
        scale = ref_norm / np.sqrt(got.size)
        err = np.abs(vals - ref_vals) / (1e-4 * scale + 1e-4 * np.abs(ref_vals) + 1e-9)
        worst = max(worst, float(err.max()))
        assert err.max() <= 1.0, (k, float(err.max()), scale)
        assert abs(np.sqrt((got.astype(np.float64) ** 2).sum()) - ref_norm) <= 1e-4 * ref_norm + 1e-9, k
    print('%s: worst gradient error %.3f of the tolerance' % (name, worst))


def test_training_step_end_to_end_loss_and_optimizer(dev, golden_dir):
    """The loop body of train.py:96-108 through the drop-in modules: loss matches the reference's, an optimizer step
    on the received gradients lowers it, and eval() returns to the inference kernels with the updated weights."""
    from pfann_b200.train import similarity_loss
    g = np.load(os.path.join(golden_dir, 'train_step.npz'))
    net, params, seed = _train_net('tiny', dev)
    x = torch.from_numpy(synth.train_case_input('tiny')).to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    losses = []
    for step in range(4):
        opt.zero_grad()
        loss = similarity_loss(net(x), params['tau'])
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert abs(losses[0] - float(g['tiny/loss'])) < 1e-4
    assert losses[-1] < losses[0]
    net.eval()
    with torch.no_grad():
        z_eval = net(x)
    net.train()
    z_train = net(x).detach()
    np.testing.assert_allclose(z_eval.cpu().numpy(), z_train.cpu().numpy(), rtol=0, atol=2e-2)   # bf16 vs fp32 kernels
    # frozen parameters get no gradient, like torch's
    for p in net.f.parameters():
        p.requires_grad_(False)
    opt.zero_grad(set_to_none=True)
    similarity_loss(net(x), params['tau']).backward()
    assert all(p.grad is None for p in net.f.parameters()) and all(p.grad is not None for p in net.g.parameters())


def test_train_step_minibatch_route_equals_the_full_batch(dev):
    """train.py:83-97: fingerprints without a graph, the loss gradient, then one backward per minibatch -- the same
    parameter gradients as the single differentiated pass (train.py:99-101)."""
    from pfann_b200.train import train_step

    class NoStep:
        def __init__(self, params):
            self.params = list(params)

        def zero_grad(self):
            for p in self.params:
                p.grad = None

        def step(self):
            pass

    x = torch.from_numpy(synth.train_case_input('tiny')).to(dev)
    grads = []
    for mb in (None, 2, 3):
        net, params, _ = _train_net('tiny', dev)
        loss = train_step(net, NoStep(net.parameters()), x, params['tau'], minibatch=mb)
        grads.append((loss.item(), [p.grad.detach().clone() for p in net.parameters()]))
    for loss, g in grads[1:]:
        assert abs(loss - grads[0][0]) < 1e-5
        for a, b in zip(g, grads[0][1]):
            assert float((a - b).norm()) <= 1e-4 * float(b.norm()) + 1e-7
