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
