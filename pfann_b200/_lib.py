"""ctypes binding of libpfann_b200.so (the C-ABI declared in include/pfann_b200.h).

Same convention as the reference's own native hook (database.py:14-32): ``cdll.LoadLibrary``, explicit
``argtypes``, a ``version()`` handshake.  There is no Python/CPU fallback: if the library cannot be
loaded, or no Blackwell GPU is present, the calls raise.
"""
import ctypes
import os
import threading
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int16, c_int32, c_int64, c_longlong, c_void_p

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'csrc', 'libpfann_b200.so')
VERSION = 20261017001

PRECISION_FP32 = 0
PRECISION_BF16 = 1

_lib = None
_lock = threading.RLock()
_ctx = {}


class PfannError(RuntimeError):
    pass


def _declare(L):
    vp = c_void_p
    L.pfann_version.restype = c_longlong
    L.pfann_last_error.restype = c_char_p
    L.pfann_ctx_create.argtypes = [c_int, POINTER(vp)]
    L.pfann_ctx_destroy.argtypes = [vp]
    L.pfann_ctx_destroy.restype = None
    L.pfann_ctx_set_stream.argtypes = [vp, vp]
    L.pfann_ctx_sync.argtypes = [vp]
    L.pfann_ctx_launches.argtypes = [vp]
    L.pfann_ctx_launches.restype = c_longlong
    L.pfann_ctx_sm_count.argtypes = [vp]
    L.pfann_ctx_profile.argtypes = [vp, c_int]
    L.pfann_ctx_profile_read.argtypes = [vp, POINTER(c_double), POINTER(c_longlong), c_int]
    L.pfann_ctx_profile_detail.argtypes = [vp, POINTER(c_double), POINTER(c_longlong), c_int]
    L.pfann_mel_create.argtypes = [vp, c_int, c_int, c_int, c_double, c_double, c_int, c_int, POINTER(vp)]
    L.pfann_mel_create_ex.argtypes = [vp, c_int, c_int, c_int, c_double, c_double, c_int, c_int, c_int, c_int, c_int, POINTER(vp)]
    L.pfann_mel_destroy.argtypes = [vp]
    L.pfann_mel_destroy.restype = None
    L.pfann_mel_forward.argtypes = [vp, vp, c_int64, vp]
    L.pfann_mel_forward_pcm16.argtypes = [vp, vp, c_int64, vp, vp, c_int64, vp]
    L.pfann_mel_n_frames.argtypes = [vp]
    L.pfann_model_create.argtypes = [vp, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(vp)]
    L.pfann_model_create_ex.argtypes = [vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_int), POINTER(vp)]
    L.pfann_model_destroy.argtypes = [vp]
    L.pfann_model_destroy.restype = None
    L.pfann_model_set_param.argtypes = [vp, c_char_p, vp, c_int64]
    L.pfann_model_finalize.argtypes = [vp, c_int]
    L.pfann_model_set_chunk.argtypes = [vp, c_int]
    L.pfann_model_forward.argtypes = [vp, vp, c_int64, c_int, vp]
    L.pfann_model_set_tap.argtypes = [vp, c_int]
    L.pfann_model_get_activation.argtypes = [vp, c_int, vp, c_int64]
    L.pfann_extract_segments.argtypes = [vp, vp, vp, c_int64, c_int, vp]
    L.pfann_extract_pcm16.argtypes = [vp, vp, vp, POINTER(c_int64), c_int, c_int, c_int, vp, POINTER(c_int32)]
    L.pfann_extract_f32.argtypes = [vp, vp, vp, POINTER(c_int64), c_int, c_int, c_int, vp, POINTER(c_int32)]
    L.pfann_pcm16_to_planar.argtypes = [vp, vp, c_int64, c_int, vp]
    L.pfann_resample_len.argtypes = [c_int64, c_int, c_int]
    L.pfann_resample_len.restype = c_int64
    L.pfann_resample_frac.argtypes = [vp, vp, c_int, c_int64, c_int, c_int, vp]
    L.pfann_mix_mono.argtypes = [vp, vp, c_int, c_int64, vp]
    L.pfann_ntxent.argtypes = [vp, vp, c_int, c_int, c_float, vp, vp]
    L.pfann_ir_conv.argtypes = [vp, vp, c_int64, c_int, vp, c_int, vp, c_int, c_int]
    L.pfann_model_train_forward.argtypes = [vp, vp, c_int64, c_int, vp]
    L.pfann_model_train_backward.argtypes = [vp, vp, c_int]
    L.pfann_model_get_grad.argtypes = [vp, c_char_p, vp, c_int64]
    L.pfann_model_train_load_param.argtypes = [vp, c_char_p, vp, c_int64]
    L.pfann_specaug_apply.argtypes = [vp, vp, vp, c_int64, c_int, c_int]
    L.pfann_snr_mix.argtypes = [vp, vp, vp, vp, c_int64, c_int, vp]
    L.pfann_count_segments.argtypes = [POINTER(c_int64), c_int, c_int, c_int]
    L.pfann_count_segments.restype = c_int64
    L.pfann_db_open.argtypes = [vp, vp, c_int64, c_int, POINTER(c_int32), c_int, c_int64, c_int64, POINTER(vp)]
    L.pfann_db_close.argtypes = [vp]
    L.pfann_db_close.restype = None
    L.pfann_db_ntotal.argtypes = [vp]
    L.pfann_db_ntotal.restype = c_int64
    L.pfann_db_set_tuning.argtypes = [vp, c_int, c_int, c_int]
    L.pfann_db_search.argtypes = [vp, vp, c_int64, c_int, vp, vp]
    L.pfann_db_seq_score.argtypes = [vp, POINTER(c_int64), c_int, POINTER(c_float), c_int, POINTER(c_int64), c_int,
                                     POINTER(c_float), c_int, c_float]
    L.pfann_db_query.argtypes = [vp, vp, POINTER(c_int64), c_int, c_int, c_int, c_float, POINTER(c_float),
                                 POINTER(c_int32), POINTER(c_float), POINTER(c_float), c_int64]
    L.pfann_topk_merge.argtypes = [vp, vp, vp, c_int, c_int64, c_int, vp, vp]
    L.pfann_db_rerank.argtypes = [vp, vp, vp, c_int, vp, c_int, c_int, c_float, vp, vp, vp]
    L.pfann_db_search_thresholds.argtypes = [vp, vp, c_int64, c_int, vp]
    L.pfann_db_search_sample_topk.argtypes = [vp, vp, c_int64, c_int, vp, vp]
    L.pfann_db_thresholds_from_topk.argtypes = [vp, vp, c_int, vp, c_int64, c_int, vp]
    L.pfann_db_search_filtered.argtypes = [vp, vp, c_int64, c_int, vp, vp, c_int]
    L.pfann_db_take_overflow.argtypes = [vp]
    L.pfann_topk_merge_keys.argtypes = [vp, vp, c_int, c_int64, c_int, vp, vp]
    L.pfann_db_rerank_packed.argtypes = [vp, vp, vp, c_int, c_int, vp, c_int, c_int, c_float, vp]
    L.pfann_best_combine.argtypes = [vp, vp, c_int, c_int, vp]
    L.pfann_db_set_sample_scale.argtypes = [vp, c_float]
    L.pfann_db_max_norm.argtypes = [vp]
    L.pfann_db_max_norm.restype = c_float
    L.pfann_db_set_max_norm.argtypes = [vp, c_float]
    # reference-compatible pair, prototypes exactly as database.py:16-29
    L.seq_score.argtypes = [c_void_p, POINTER(c_int64), c_int, POINTER(c_float), c_int, POINTER(c_int64), c_int,
                            POINTER(c_float), c_int, c_float]
    L.seq_score.restype = c_int
    L.version.restype = c_int64


def lib():
    """Load (building first if the in-tree .so is absent) and version-check the native library."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                from . import build as _build
                _build.build()
            L = ctypes.CDLL(LIB_PATH)
            _declare(L)
            if L.pfann_version() != VERSION or L.version() != 20220625002:
                raise PfannError('libpfann_b200.so: wrong version, please rebuild (python -m pfann_b200.build --force)')
            _lib = L
    return _lib


def check(rc, what=''):
    if rc < 0:
        msg = lib().pfann_last_error()
        raise PfannError('%s failed (%d): %s' % (what or 'libpfann_b200 call', rc, msg.decode() if msg else ''))
    return rc


def ctx(device=0):
    """Per-device context handle (created once per process and device)."""
    L = lib()
    with _lock:
        if device not in _ctx:
            h = c_void_p()
            check(L.pfann_ctx_create(int(device), ctypes.byref(h)), 'pfann_ctx_create')
            _ctx[device] = h
    return _ctx[device]


def use_torch_stream(device):
    """Make the context enqueue on torch's current stream of `device` (torch = plumbing: memory + streams)."""
    import torch
    h = ctx(device)
    s = torch.cuda.current_stream(device).cuda_stream
    check(lib().pfann_ctx_set_stream(h, c_void_p(s)), 'pfann_ctx_set_stream')
    return h


def launches(device=0):
    return int(lib().pfann_ctx_launches(ctx(device)))


KERNEL_CLASSES = ['mel', 'conv_tc', 'conv_cc', 'layernorm', 'head', 'knn_scan', 'knn_select', 'rerank', 'misc']


def profile(device, enable):
    check(lib().pfann_ctx_profile(ctx(device), int(enable)), 'pfann_ctx_profile')


def profile_read(device=0):
    """{class: (milliseconds, launches)} accumulated since the last read (synchronises the stream)."""
    n = len(KERNEL_CLASSES)
    ms = (c_double * n)()
    cnt = (c_longlong * n)()
    check(lib().pfann_ctx_profile_read(ctx(device), ms, cnt, n), 'pfann_ctx_profile_read')
    return {k: (ms[i], int(cnt[i])) for i, k in enumerate(KERNEL_CLASSES)}


def profile_detail(device=0):
    """Per-kernel breakdown of the record consumed by the last profile_read(): {name: (ms, launches)}."""
    n = 48
    ms = (c_double * n)()
    cnt = (c_longlong * n)()
    check(lib().pfann_ctx_profile_detail(ctx(device), ms, cnt, n), 'pfann_ctx_profile_detail')
    names = ['conv%d' % i for i in range(16)] + ['ln%d' % i for i in range(16)] + ['mel', 'head', 'l0_moments', 'knn_scan_sample', 'knn_scan_full', 'knn_kth', 'knn_select', 'rerank']
    return {k: (ms[i], int(cnt[i])) for i, k in enumerate(names) if cnt[i]}


def ptr(a):
    """Raw pointer of a torch tensor (CPU or CUDA) or numpy array; the buffer must be contiguous."""
    if a is None:
        return c_void_p(None)
    if isinstance(a, np.ndarray):
        assert a.flags['C_CONTIGUOUS']
        return c_void_p(a.ctypes.data)
    assert a.is_contiguous()
    return c_void_p(a.data_ptr())


def device_index(t):
    """CUDA device ordinal for a torch tensor; CPU tensors use the current CUDA device."""
    import torch
    if t.is_cuda:
        return t.device.index if t.device.index is not None else torch.cuda.current_device()
    if not torch.cuda.is_available():
        raise PfannError('pfann_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
    return torch.cuda.current_device()
