"""Drop-in for the reference's ``datautil/melspec.py``: same class, same factory, same forward contract,
computed by the fused sm_100a kernel behind ``pfann_mel_forward`` (include/pfann_b200.h).

    mel = build_mel_spec_layer(params).to(device)       # builder.py:68, matcher.py:77
    g = mel(batch)                                      # [..., n] -> [..., n_mels, T]   melspec.py:33-50
"""
import ctypes

import torch

from .. import _lib


class MelSpec(torch.nn.Module):
    """Mirror of MelSpec (datautil/melspec.py:4-50) with all of its options: naf_mode (magnitude spectrum, zero
    padding, slaney mel scale + normalisation, + 0.06), mel_log ('log' | 'log10' | anything else = none, as
    melspec.py:43-46 falls through) and spec_norm ('l2' | 'max')."""

    def __init__(self, sample_rate=8000, stft_n=1024, stft_hop=256, f_min=300, f_max=4000, n_mels=256,
                 naf_mode=False, mel_log='log', spec_norm='l2'):
        super(MelSpec, self).__init__()
        self.sample_rate, self.stft_n, self.stft_hop = sample_rate, stft_n, stft_hop
        self.f_min, self.f_max, self.n_mels = f_min, f_max, n_mels
        self.naf_mode, self.mel_log, self.spec_norm = naf_mode, mel_log, spec_norm
        self._plans = {}

    def _plan(self, device, seg_len):
        key = (device, seg_len)
        if key not in self._plans:
            h = ctypes.c_void_p()
            log_mode = {'log': 1, 'log10': 2}.get(self.mel_log, 0)
            _lib.check(_lib.lib().pfann_mel_create_ex(_lib.ctx(device), self.sample_rate, self.stft_n, self.stft_hop,
                                                      float(self.f_min), float(self.f_max), self.n_mels, seg_len,
                                                      int(bool(self.naf_mode)), log_mode,
                                                      int(self.spec_norm == 'max'), ctypes.byref(h)),
                       'pfann_mel_create_ex')
            self._plans[key] = h
        return self._plans[key]

    def plan_handle(self, device, seg_len):
        return self._plan(device, seg_len)

    def forward(self, x):
        lead, n = x.shape[:-1], x.shape[-1]
        dev = _lib.device_index(x)
        h = self._plan(dev, n)
        _lib.use_torch_stream(dev)
        xf = x.reshape(-1, n).to(torch.float32).contiguous()
        T = 1 + n // self.stft_hop
        out = torch.empty((xf.shape[0], self.n_mels, T), dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().pfann_mel_forward(h, _lib.ptr(xf), xf.shape[0], _lib.ptr(out)), 'pfann_mel_forward')
        return out.reshape(*lead, self.n_mels, T)

    def __del__(self):
        try:
            for h in self._plans.values():
                _lib.lib().pfann_mel_destroy(h)
        except Exception:
            pass


def build_mel_spec_layer(params):
    """datautil/melspec.py:52-63."""
    return MelSpec(
        sample_rate=params['sample_rate'],
        stft_n=params['stft_n'],
        stft_hop=params['stft_hop'],
        f_min=params['f_min'],
        f_max=params['f_max'],
        n_mels=params['n_mels'],
        naf_mode=params.get('naf_mode', False),
        mel_log=params.get('mel_log', 'log'),
        spec_norm=params.get('spec_norm', 'l2'),
    )
