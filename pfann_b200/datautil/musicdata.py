"""The segmenter boundary of the hot path (datautil/musicdata.py:72-88 of the reference).

Decoding and resampling are out of scope (I/O-bound CPU work, and
``julius.ResampleFrac`` is the identity at the model's own rate): this module reads 16-bit PCM WAV files that
are already at ``params['sample_rate']`` and hands them to the GPU, where zero-padding, framing and per-row
mean removal happen inside the mel kernel (``pfann_mel_forward_pcm16``).  Anything else is reported like the
reference reports an unreadable file (musicdata.py:95-101): zero segments, never an abort.
"""
import wave

import numpy as np


class UnsupportedAudio(Exception):
    pass


def read_wav_pcm16(path, sample_rate):
    """Returns ('pcm16', int16[n]) for mono files, or ('float', float32[n]) for multi-channel files mixed down
    the way musicdata.py:48,72-80 does (scale by 1/32768, fake-stereo check, mean over channels)."""
    with wave.open(path, 'rb') as w:
        if w.getsampwidth() != 2:
            raise UnsupportedAudio('%s: only 16-bit PCM WAV is supported (audio.py:137-138)' % path)
        if w.getframerate() != sample_rate:
            raise UnsupportedAudio('%s: %d Hz; resample to %d Hz first (ingest/resampling is out of scope)'
                                   % (path, w.getframerate(), sample_rate))
        nch = w.getnchannels()
        data = np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16)
    if nch == 1:
        return 'pcm16', data
    x = np.multiply(data.reshape(-1, nch), 1 / 32768, dtype=np.float32).T     # musicdata.py:47-48
    if nch == 2:                                                              # musicdata.py:74-79
        pow1 = ((x[0] - x[1]) ** 2).mean()
        pow2 = ((x[0] + x[1]) ** 2).mean()
        if pow1 > pow2 * 1000:
            x[1] *= -1
    return 'float', x.mean(axis=0).astype(np.float32)                         # musicdata.py:80


def frame_float(wav, seg, hop):
    """musicdata.py:82-88 on the host, for the (rare) multi-channel inputs."""
    if wav.shape[0] < seg:
        wav = np.pad(wav, (0, seg - wav.shape[0]))
    n_seg = (wav.shape[0] - seg) // hop + 1
    rows = np.lib.stride_tricks.as_strided(wav, (n_seg, seg), (hop * wav.itemsize, wav.itemsize)).copy()
    return rows - rows.mean(axis=1, dtype=np.float32, keepdims=True)
