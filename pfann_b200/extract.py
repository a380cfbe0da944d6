"""Host side of the builder / matcher inner loop (builder.py:75-103, matcher.py:85-131) on top of the C-ABI:
PCM (or framed rows) in, fingerprints out, with the segmenter tail (musicdata.py:82-88), the log-mel and the
network fused on the device.  This is the public call ``bench.py`` measures end to end."""
import ctypes
from ctypes import POINTER, c_int32, c_int64

import numpy as np
import torch

from . import _lib, synth
from .datautil.melspec import build_mel_spec_layer
from .model import FpNetwork


class Extractor:
    """mel + model pair for one JSON config (the `params` dict of builder.py:35-51)."""

    def __init__(self, params, state_dict=None, device=None, precision=None, chunk=None):
        if not torch.cuda.is_available():
            raise _lib.PfannError('pfann_b200.Extractor needs a CUDA device (sm_100a); there is no CPU fallback')
        self.params = params
        self.device = torch.device('cuda', torch.cuda.current_device() if device is None else int(device))
        d, h, u, F, T = synth.model_dims(params)
        self.d = d
        self.seg_len = int(params['segment_size'] * params['sample_rate'])
        self.hop = int(params['hop_size'] * params['sample_rate'])
        mp = dict(params['model'])
        if precision:
            mp['b200_precision'] = precision
        if chunk:
            mp['b200_chunk'] = int(chunk)
        self.model = FpNetwork(d, h, u, F, T, mp).to(self.device)
        if state_dict is not None:
            self.model.load_state_dict({k: torch.as_tensor(v) for k, v in state_dict.items()})
        self.model.eval()
        for p in self.model.parameters():
            p.requires_grad = False                                   # builder.py:60-62
        self.mel = build_mel_spec_layer(params).to(self.device)

    def _handles(self):
        dev = self.device.index
        return self.mel.plan_handle(dev, self.seg_len), self.model.native_handle(dev), dev

    def count_segments(self, clip_off, frame_shift_mul=1):
        clip_off = np.ascontiguousarray(clip_off, dtype=np.int64)
        return int(_lib.lib().pfann_count_segments(clip_off.ctypes.data_as(POINTER(c_int64)), len(clip_off) - 1,
                                                   self.seg_len, self.hop // frame_shift_mul))

    def extract_pcm16(self, pcm, clip_off, frame_shift_mul=1, norm=True, out=None):
        """pcm: int16 mono at the model rate, clips back to back (numpy / CPU tensor / CUDA tensor);
        clip_off: [n_clips+1] sample offsets.  Returns (z [n_seg, d] on pcm's side of the bus, seg_counts)."""
        clip_off = np.ascontiguousarray(clip_off, dtype=np.int64)
        n_clips = len(clip_off) - 1
        n_seg = self.count_segments(clip_off, frame_shift_mul)
        on_dev = isinstance(pcm, torch.Tensor) and pcm.is_cuda
        if out is None:
            out = (torch.empty((n_seg, self.d), dtype=torch.float32, device=self.device) if on_dev
                   else np.empty((n_seg, self.d), np.float32))
        counts = np.empty(n_clips, np.int32)
        hm, hn, dev = self._handles()
        _lib.use_torch_stream(dev)
        _lib.check(_lib.lib().pfann_extract_pcm16(hm, hn, _lib.ptr(pcm), clip_off.ctypes.data_as(POINTER(c_int64)),
                                                  n_clips, self.hop // frame_shift_mul, int(bool(norm)),
                                                  _lib.ptr(out), counts.ctypes.data_as(POINTER(c_int32))),
                   'pfann_extract_pcm16')
        return out, counts

    def ingest_wav(self, pcm, rate):
        """Decoded 16-bit PCM of any channel count and rate (int16 [n_frames, nch]) -> fp32 mono at the model rate
        on the GPU: musicdata.py:44-80 (scale, julius-style fractional resampling, mono mix with the fake-stereo
        rule).  Returns a 1-D CUDA tensor."""
        L, dev = _lib.lib(), self.device.index
        pcm = np.array(pcm, dtype=np.int16, order='C')           # own, writable copy (WAV readers hand out views)
        if pcm.ndim == 1:
            pcm = pcm[:, None]
        n, nch = pcm.shape
        h = _lib.use_torch_stream(dev)
        src = torch.from_numpy(pcm).to(self.device)
        planar = torch.empty((nch, n), dtype=torch.float32, device=self.device)
        _lib.check(L.pfann_pcm16_to_planar(h, _lib.ptr(src), n, nch, _lib.ptr(planar)), 'pfann_pcm16_to_planar')
        sr = self.params['sample_rate']
        if rate != sr:
            n_out = int(L.pfann_resample_len(n, int(rate), int(sr)))
            res = torch.empty((nch, n_out), dtype=torch.float32, device=self.device)
            _lib.check(L.pfann_resample_frac(h, _lib.ptr(planar), nch, n, int(rate), int(sr), _lib.ptr(res)),
                       'pfann_resample_frac')
            planar, n = res, n_out
        mono = torch.empty(n, dtype=torch.float32, device=self.device)
        _lib.check(L.pfann_mix_mono(h, _lib.ptr(planar), nch, n, _lib.ptr(mono)), 'pfann_mix_mono')
        return mono

    def extract_f32(self, wav, clip_off, frame_shift_mul=1, norm=True):
        """wav: fp32 mono at the model rate, clips back to back, ON THE GPU (e.g. from ingest_wav); like extract_pcm16."""
        clip_off = np.ascontiguousarray(clip_off, dtype=np.int64)
        n_clips = len(clip_off) - 1
        n_seg = self.count_segments(clip_off, frame_shift_mul)
        out = torch.empty((n_seg, self.d), dtype=torch.float32, device=self.device)
        counts = np.empty(n_clips, np.int32)
        hm, hn, dev = self._handles()
        _lib.use_torch_stream(dev)
        _lib.check(_lib.lib().pfann_extract_f32(hm, hn, _lib.ptr(wav.contiguous()), clip_off.ctypes.data_as(POINTER(c_int64)),
                                                n_clips, self.hop // frame_shift_mul, int(bool(norm)), _lib.ptr(out),
                                                counts.ctypes.data_as(POINTER(c_int32))), 'pfann_extract_f32')
        return out, counts

    def extract_wavs(self, wavs, frame_shift_mul=1, norm=True):
        """wavs: list of (int16 [n_frames, nch], rate) as read from WAV files -> (z numpy [n_seg, d], seg_counts):
        GPU ingest of every clip, then ONE fused framing + mel + network call over all of them."""
        monos = [self.ingest_wav(p, r) for p, r in wavs]
        off = np.concatenate([[0], np.cumsum([m.shape[0] for m in monos])]).astype(np.int64)
        wav = torch.cat(monos) if monos else torch.zeros(0, dtype=torch.float32, device=self.device)
        z, counts = self.extract_f32(wav, off, frame_shift_mul, norm)
        return z.cpu().numpy(), counts

    def extract_segments(self, rows, norm=True):
        """rows [B, seg_len] fp32 as MusicDataset yields them (musicdata.py:87-88) -> z [B, d]."""
        on_dev = isinstance(rows, torch.Tensor) and rows.is_cuda
        B = rows.shape[0]
        out = (torch.empty((B, self.d), dtype=torch.float32, device=self.device) if on_dev
               else np.empty((B, self.d), np.float32))
        if isinstance(rows, np.ndarray):
            rows = np.ascontiguousarray(rows, dtype=np.float32)
        hm, hn, dev = self._handles()
        _lib.use_torch_stream(dev)
        _lib.check(_lib.lib().pfann_extract_segments(hm, hn, _lib.ptr(rows), B, int(bool(norm)), _lib.ptr(out)),
                   'pfann_extract_segments')
        return out
