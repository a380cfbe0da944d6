"""GPU probe: per-kernel event timing of one extraction chunk sequence (conv idx, LayerNorm kernels, mel, head).

    python tools/conv_probe.py [clips] [chunk]

Environment switches of the library (PFANN_B200_NO_FUSED_LN, PFANN_B200_NO_BRES, ...) apply as usual.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pfann_b200 import _lib, synth  # noqa: E402
from pfann_b200.extract import Extractor  # noqa: E402

clips = int(sys.argv[1]) if len(sys.argv) > 1 else 700
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
params = synth.read_config('default')
ex = Extractor(params, synth.make_state_dict(params, seed=11), device=0, precision='bf16', chunk=chunk)
clip = 240000
pcm = torch.randint(-8000, 8000, (clips * clip,), dtype=torch.int16, device='cuda')
off = np.arange(clips + 1, dtype=np.int64) * clip
for _ in range(2):
    ex.extract_pcm16(pcm, off)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ex.extract_pcm16(pcm, off)
e1.record()
torch.cuda.synchronize()
nseg = clips * 59
print('segments %d  total %.2f ms  -> %.0f fingerprints/s (unprofiled)' % (nseg, e0.elapsed_time(e1), nseg / e0.elapsed_time(e1) * 1e3))
_lib.profile(0, True)
ex.extract_pcm16(pcm, off)
p = _lib.profile_read(0)
d = _lib.profile_detail(0)
_lib.profile(0, False)
nchunks = (nseg + chunk - 1) // chunk
print('classes (ms): ' + ', '.join('%s %.2f' % (k, v[0]) for k, v in p.items() if v[1]))
tot = sum(v[0] for v in d.values())
print('per chunk of %d segments (us), %d chunks:' % (chunk, nchunks))
for k, v in d.items():
    print('  %-11s %9.1f us  (%4.1f %%)  launches/chunk %.1f' % (k, v[0] * 1e3 / nchunks, 100 * v[0] / tot, v[1] / nchunks))
print('  sum %.1f us/chunk' % (tot * 1e3 / nchunks))
