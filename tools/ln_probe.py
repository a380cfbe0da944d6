"""GPU probe: cycle accounting inside the fused conv+LayerNorm kernel, per convolution."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pfann_b200 import _lib, synth  # noqa: E402
from pfann_b200.extract import Extractor  # noqa: E402

prof = torch.zeros((16, 148, 8), dtype=torch.int64, device='cuda')
os.environ['PFANN_LN_PROF_PTR'] = str(prof.data_ptr())
params = synth.read_config('default')
ex = Extractor(params, synth.make_state_dict(params, seed=11), device=0, precision='bf16', chunk=4096)
clips, clip = 140, 240000
pcm = torch.randint(-8000, 8000, (clips * clip,), dtype=torch.int16, device='cuda')
off = np.arange(clips + 1, dtype=np.int64) * clip
ex.extract_pcm16(pcm, off)
torch.cuda.synchronize()
prof.zero_()
_lib.profile(0, True)
ex.extract_pcm16(pcm, off)
p = _lib.profile_read(0)
print({k: round(v[0], 3) for k, v in p.items() if v[1]})
pr = prof.cpu().numpy().astype(np.float64)
names = ['wait MMA', 'pass1', 'wait stats', 'pass2', '| MMA: wait slot', 'wait TMA', '| TMA: wait stage']
for idx in range(1, 15):
    g = pr[idx, :, 4].sum() / 1.0
    if g == 0:
        continue
    per = [pr[idx, :, j].sum() / 16.0 / g for j in range(4)]   # cycles per tile per CTA (avg of 16 warps)
    per += [pr[idx, :, j].sum() / g for j in (5, 6, 7)]          # single-thread roles
    print('conv %2d: tiles/CTA %.0f | cycles per tile: %s' % (idx, g / 148, ', '.join('%s %.0f' % (n, v) for n, v in zip(names, per))))
