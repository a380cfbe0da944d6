"""GPU probe: cycle accounting inside front_tc_kernel (average over the 16 worker warps, per tile and CTA)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pfann_b200 import _lib, synth  # noqa: E402
from pfann_b200.extract import Extractor  # noqa: E402

prof = torch.zeros((148, 16), dtype=torch.int64, device='cuda')
os.environ['PFANN_FRONT_PROF_PTR'] = str(prof.data_ptr())
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
d = _lib.profile_detail(0)
print({k: round(v[0], 3) for k, v in d.items() if k in ('conv0', 'conv1', 'conv2', 'mel')})
pr = prof.cpu().numpy().astype(np.float64)
tiles = pr[:, 7].sum()
names = ['wait staged mel', 'wait row slots', 'produce', 'wait stats', 'pass2', 'wait MMA', 'pass1']
per = [pr[:, j].sum() / (16.0 if j < 3 else 8.0) / tiles for j in range(7)]   # 16 producer warps, 8 epilogue warps
print('tiles/CTA %.0f | cycles per tile: producers: %s (sum %.0f) | epilogue: %s (sum %.0f)' % (
    tiles / 128, ', '.join('%s %.0f' % (n, v) for n, v in zip(names[:3], per[:3])), sum(per[:3]),
    ', '.join('%s %.0f' % (n, v) for n, v in zip(names[3:], per[3:])), sum(per[3:])))
slow = pr[:, 3] / 8.0 / np.maximum(pr[:, 7], 1)
print('wait-stats cycles per tile by CTA: min %.0f median %.0f max %.0f' % (slow[:128].min(), np.median(slow[:128]), slow[:128].max()))
order = np.argsort(slow[:128])
print('slowest CTAs (least waiting): ' + ', '.join('fo %d sm %d wait %.0f prod %.0f' % (i, pr[i, 9], slow[i], pr[i, 2] / 16 / pr[i, 7]) for i in order[:10]))
print('fastest CTAs (most waiting): ' + ', '.join('fo %d sm %d wait %.0f prod %.0f' % (i, pr[i, 9], slow[i], pr[i, 2] / 16 / pr[i, 7]) for i in order[-6:]))
print('MMA thread per tile: total %.0f, wait TMEM slot %.0f, wait rows %.0f' % ((pr[:128, 8] / pr[:128, 7]).mean(), (pr[:128, 10] / pr[:128, 7]).mean(), (pr[:128, 11] / pr[:128, 7]).mean()))
sms = pr[:128, 9].astype(int)
tpc = sms // 2
both = np.array([np.sum(tpc == t) for t in tpc])
print('CTAs alone on their TPC: %d, sharing: %d; mean produce cycles alone %.0f / sharing %.0f' % (
    (both == 1).sum(), (both == 2).sum(), (pr[:128, 2] / 16 / pr[:128, 7])[both == 1].mean() if (both == 1).any() else 0,
    (pr[:128, 2] / 16 / pr[:128, 7])[both == 2].mean()))
