"""The segmenter boundary of the hot path (datautil/musicdata.py:29-88 of the reference).

Decoding compressed audio is out of scope (ffmpeg, I/O-bound CPU work): this module reads
16-bit PCM WAV files.  Everything after the decoder runs on the GPU:

  * mono files already at ``params['sample_rate']`` are handed over as int16 PCM; zero-padding, framing and per-row
    mean removal happen inside the mel kernel (``pfann_extract_pcm16``);
  * any other channel count / sample rate goes through the GPU ingest (``Extractor.extract_wavs``): planar fp32,
    fractional resampling (the julius.ResampleFrac algorithm of musicdata.py:29), mono mix with the fake-stereo rule
    of musicdata.py:72-80, then the same fused framing + mel + network.

Unreadable files are reported like the reference reports them (musicdata.py:95-101): zero segments, never an abort.
"""
import wave

import numpy as np


class UnsupportedAudio(Exception):
    pass


def read_wav(path):
    """(int16[n_frames, nch], sample_rate) of a 16-bit PCM WAV file (audio.py:130-150)."""
    with wave.open(path, 'rb') as w:
        if w.getsampwidth() != 2:
            raise UnsupportedAudio('%s: only 16-bit PCM WAV is supported (audio.py:137-138)' % path)
        nch, rate = w.getnchannels(), w.getframerate()
        data = np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16)
    return data.reshape(-1, nch), rate


def read_wav_pcm16(path, sample_rate):
    """('pcm16', int16[n]) for mono files at the model rate (the fast path), else ('wav', (int16[n, nch], rate)) for
    the GPU ingest."""
    data, rate = read_wav(path)
    if data.shape[1] == 1 and rate == sample_rate:
        return 'pcm16', np.ascontiguousarray(data[:, 0])
    return 'wav', (data, rate)


def mix_float_host(data, nch):
    """musicdata.py:47-48,72-80 on the host (numpy): kept as the readable statement of the rule the GPU kernel
    implements (tests compare the two); the command lines use the GPU ingest."""
    x = np.multiply(data.reshape(-1, nch), 1 / 32768, dtype=np.float32).T     # musicdata.py:47-48
    if nch == 2:                                                              # musicdata.py:74-79
        pow1 = ((x[0] - x[1]) ** 2).mean()
        pow2 = ((x[0] + x[1]) ** 2).mean()
        if pow1 > pow2 * 1000:
            x[1] *= -1
    return x.mean(axis=0).astype(np.float32)                                  # musicdata.py:80


def frame_float(wav, seg, hop):
    """musicdata.py:82-88 on the host (tests only; the product frames inside the mel kernel)."""
    if wav.shape[0] < seg:
        wav = np.pad(wav, (0, seg - wav.shape[0]))
    n_seg = (wav.shape[0] - seg) // hop + 1
    rows = np.lib.stride_tricks.as_strided(wav, (n_seg, seg), (hop * wav.itemsize, wav.itemsize)).copy()
    return rows - rows.mean(axis=1, dtype=np.float32, keepdims=True)
