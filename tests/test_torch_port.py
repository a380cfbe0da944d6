"""CPU: the torch-CPU port used as the bench's CPU baseline reproduces the reference's golden outputs."""
import os

import numpy as np
import torch

from oracle.torch_port import TorchPort
from pfann_b200 import synth


def test_torch_port_vs_reference_golden(golden_dir):
    x = np.concatenate([synth.synth_segments(3, seed=1), np.zeros((1, 8000), np.float32)])
    gm = np.load(os.path.join(golden_dir, 'mel_default.npz'))['mel']
    for name in ('tiny', 'n640d64', 'default'):
        params = synth.read_config(name)
        g = np.load(os.path.join(golden_dir, 'enc_%s.npz' % name))
        port = TorchPort(params, synth.make_state_dict(params, seed=int(g['seed'])))
        with torch.no_grad():
            mel = port.mel(torch.from_numpy(x))
            z = port.encoder(torch.from_numpy(gm)).numpy()
        np.testing.assert_allclose(mel.numpy(), gm, rtol=0, atol=2e-4)
        np.testing.assert_allclose(z, g['z'], rtol=0, atol=1e-5)
