"""torch-CPU restatement of stages 1+2 (the reference's own execution model: PyTorch library ops on the host).
TEST INFRASTRUCTURE / CPU BASELINE ONLY -- imported solely by tests/ and by bench.py's cpu_baseline and
``--impl reference`` legs; never by the product package.

/root/reference cannot travel to the GPU box, so the baseline the bench times there is this port: the same
sequence of library calls the reference makes (torch.stft -> mel matmul -> log; per layer ZeroPad2d ->
Conv2d -> LayerNorm -> ReLU twice; grouped Conv1d head), written functionally from model.py:54-73,122-130
and datautil/melspec.py:33-50, and checked against the golden vectors in tests/test_torch_port.py.
"""
import math

import torch
import torch.nn.functional as F


def mel_fbanks(params):
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale='htk') (melspec.py:19-31)."""
    n_freqs = params['stft_n'] // 2 + 1
    n_mels = params['n_mels']
    all_freqs = torch.linspace(0, params['sample_rate'] // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + params['f_min'] / 700.0)
    m_max = 2595.0 * math.log10(1.0 + params['f_max'] / 700.0)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up))


class TorchPort:
    def __init__(self, params, state_dict):
        self.p = params
        self.sd = {k: torch.as_tensor(v) for k, v in state_dict.items()}
        self.fb = mel_fbanks(params)
        self.win = torch.hann_window(params['stft_n'])
        self.fuller = bool(params['model'].get('fuller', False))

    def mel(self, x):
        p = self.p
        x = F.normalize(x, p=2, dim=-1)                                            # melspec.py:36
        spec = torch.stft(x, p['stft_n'], hop_length=p['stft_hop'], win_length=p['stft_n'], window=self.win,
                          center=True, pad_mode='reflect', normalized=False, onesided=True, return_complex=True)
        spec = spec.abs().pow(2.0)
        mel = torch.matmul(spec.transpose(-1, -2), self.fb).transpose(-1, -2)
        return torch.log(mel + 1e-8)                                               # melspec.py:41,46

    def encoder(self, x, norm=True):
        sd = self.sd
        x = x.unsqueeze(1)
        for l in range(8):
            g = lambda k: sd['f.convs.%d.%s' % (l, k)]
            T, Fq = x.shape[3], x.shape[2]
            pad = (T - 1) // 2 * 2 + 3 - T
            x = F.pad(x, (pad // 2, pad - pad // 2, 0, 0))                         # model.py:18-19
            x = F.conv2d(x, g('conv1.weight'), g('conv1.bias'), stride=(1, 2))
            x = F.relu(F.layer_norm(x, x.shape[1:], g('ln1.weight'), g('ln1.bias'), 1e-5))
            pad = (Fq - 1) // 2 * 2 + 3 - Fq
            x = F.pad(x, (0, 0, pad // 2, pad - pad // 2))                         # model.py:24-25
            x = F.conv2d(x, g('conv2.weight'), g('conv2.bias'), stride=(2, 1),
                         groups=1 if self.fuller else x.shape[1])
            x = F.relu(F.layer_norm(x, x.shape[1:], g('ln2.weight'), g('ln2.bias'), 1e-5))
        m = self.p['model']
        x = x.reshape(-1, m['h'], 1)                                                # model.py:123
        x = F.elu(F.conv1d(x, sd['g.linear1.weight'], sd['g.linear1.bias'], groups=m['d']))
        x = F.conv1d(x, sd['g.linear2.weight'], sd['g.linear2.bias'], groups=m['d']).reshape(-1, m['d'])
        return F.normalize(x, p=2.0) if norm else x

    @torch.no_grad()
    def extract(self, rows, batch=32):
        """builder.py:88-99: split 32 -> mel -> model."""
        out = []
        for b in torch.split(rows, batch):
            out.append(self.encoder(self.mel(b)))
        return torch.cat(out)


def similarity_loss(y, tau):
    """train.py:41-52 with library ops (masked log-softmax instead of the per-row python loop; same numbers)."""
    n = y.shape[0]
    a = torch.matmul(y, y.T) / tau
    a = a.masked_fill(torch.eye(n, dtype=torch.bool, device=y.device), float('-inf'))   # the row's own entry is left out
    ls = F.log_softmax(a, dim=1)
    idx = torch.arange(n, device=y.device)
    return ls[idx, idx ^ 1].sum() / -n


def train_step_cpu(port, b, tau, lr=1e-4):
    """CPU baseline of one training step with the reference's library calls: dataset_v2.py:152-169 on the batch
    b = {orig, aug, noise, snr, air, mic} (SNR mix noise.py:96-109, rfft * H * H -> irfft, interleave) -> mel -> encoder
    under autograd -> NT-Xent -> backward -> Adam (train.py:78-104 without SpecAugment's draws).  Returns the loss."""
    params = [p.requires_grad_(True) for p in port.sd.values()]
    opt = torch.optim.Adam(params, lr=lr)
    with torch.no_grad():
        x, noise = b['aug'], b['noise']
        vx = x.pow(2).mean(dim=1).clamp_min(1e-12).sqrt()
        vn = noise.pow(2).mean(dim=1).clamp_min(1e-12).sqrt()
        x = x + (vx / vn * torch.pow(10.0, -b['snr'] / 20.0))[:, None] * noise
        n = 1024
        while n < x.shape[1] + b['air'].shape[1] + b['mic'].shape[1]:
            n *= 2
        spec = torch.fft.rfft(x, n) * torch.fft.rfft(b['air'], n) * torch.fft.rfft(b['mic'], n)
        x = torch.fft.irfft(spec, n)[..., :x.shape[1]]
        g = port.mel(torch.stack([b['orig'], x], dim=1).flatten(0, 1))
    opt.zero_grad()
    loss = similarity_loss(port.encoder(g), tau)
    loss.backward()
    opt.step()
    return float(loss.detach())
