#!/usr/bin/env python
"""Training-step timing (BASELINE configs[4]: train.py, NT-Xent, configs/n640d64.json, augment + melspec + encoder).

    python tools/train_bench.py [--steps 10 --warmup 3 --batch 640 --config n640d64] [--check]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/train_bench.py --gpus N ...

A step = the loop body of train.py:78-104 on a resident batch of PCM segments: SNR noise mix (datautil/noise.py:96-109)
-> room + microphone impulse responses (dataset_v2.py:157-163) -> log-mel -> SpecAugment -> encoder forward -> NT-Xent over the GLOBAL batch -> backward -> gradient sum over the
ranks (one all-reduce) -> Adam.  --batch is per rank (weak scaling); the loss always spans world x batch rows.
Prints one JSON line (rank 0).  --check: at N > 1, rank 0 repeats the first step alone on the whole global batch and
compares the summed gradients of the ranks with it (they are the same function)."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pfann_b200 import _lib, synth  # noqa: E402
from pfann_b200.datautil.melspec import build_mel_spec_layer  # noqa: E402
from pfann_b200.model import FpNetwork  # noqa: E402
from pfann_b200.train import SpecAugment, add_noises, apply_ir, similarity_loss, train_step  # noqa: E402


def make_batch(params, n, seed, dev):
    """What MusicSegmentDataset.__getitem__ holds before augmenting (dataset_v2.py:139-150): x_orig, x_aug [n/2][segn]
    (a clip and a time-shifted copy), plus the noise rows, SNRs (noise.py:96-109) and the room (1 s) and microphone
    (0.5 s) impulse responses picked for the augmented half."""
    sr = params['sample_rate']
    segn = int(params['segment_size'] * sr)
    rng = np.random.Generator(np.random.PCG64(seed))
    clips = synth.synth_segments(n // 2, seed=seed, seg=segn + 800)
    noise = rng.standard_normal((n // 2, segn)).astype(np.float32)
    snr = rng.uniform(0, 10, n // 2).astype(np.float32)
    air = (rng.standard_normal((n // 2, sr)) * np.exp(-np.arange(sr) / 900.0)).astype(np.float32)
    mic = (rng.standard_normal((n // 2, sr // 2)) * np.exp(-np.arange(sr // 2) / 60.0)).astype(np.float32)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    return {'orig': t(clips[:, :segn]), 'aug': t(clips[:, 800:]), 'noise': t(noise), 'snr': t(snr), 'air': t(air),
            'mic': t(mic)}


def augment(b):
    """dataset_v2.py:152-169: noise, then impulse responses, on the augmented half; rows [orig_0, aug_0, orig_1, ...]."""
    aug = apply_ir(add_noises(b['aug'], b['noise'], b['snr']), [b['air'], b['mic']])
    return torch.stack([b['orig'], aug], dim=1).flatten(0, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batch', type=int, default=640)
    ap.add_argument('--config', default='n640d64')
    ap.add_argument('--check', action='store_true')
    ap.add_argument('--cpu-baseline', type=int, default=0, metavar='ROWS',
                    help='also time one step of the torch-CPU restatement (oracle/torch_port.py) on ROWS rows')
    ap.add_argument('--profile', action='store_true', help='one extra step with a CUDA-event pair around every kernel')
    args = ap.parse_args()
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    params = synth.read_config(args.config)
    d, h, u, F, T = synth.model_dims(params)
    sd = {k: torch.from_numpy(v) for k, v in synth.make_state_dict(params, seed=3).items()}

    def new_model():
        net = FpNetwork(d, h, u, F, T, params['model']).to(dev)
        net.load_state_dict(sd)
        return net.train()

    model = new_model()
    mel = build_mel_spec_layer(params)
    specaug = SpecAugment(params)
    opt = torch.optim.Adam(model.parameters(), lr=params.get('lr', 1e-4))
    tau = params.get('tau', 0.05)
    batch = make_batch(params, args.batch, 100 + rank, dev)
    ctxh = _lib.ctx(local)
    L = _lib.lib()

    def step():
        torch.manual_seed(1)      # same SpecAugment draws on every rank and every step: a fixed workload
        g = mel(augment(batch))
        return train_step(model, opt, g, tau, specaug=specaug)

    check = None
    if args.check and world > 1:
        # the ranks' summed gradients == one GPU on the concatenated batch
        torch.manual_seed(1)
        g_local = specaug.augment(mel(augment(batch)))
        from pfann_b200.train import allreduce_gradients, similarity_loss_gathered
        model.zero_grad()
        similarity_loss_gathered(model(g_local), tau).backward()
        allreduce_gradients(list(model.parameters()))
        parts = [torch.empty_like(g_local) for _ in range(world)]
        dist.all_gather(parts, g_local)
        if rank == 0:
            solo = new_model()
            similarity_loss(solo(torch.cat(parts)), tau).backward()
            num = sum(float(((a.grad - b.grad) ** 2).sum()) for a, b in zip(model.parameters(), solo.parameters()))
            den = sum(float((b.grad ** 2).sum()) for b in solo.parameters())
            check = {'rel_l2_error_of_summed_gradients': (num / den) ** 0.5}
            del solo
        model.zero_grad()

    for _ in range(args.warmup):
        loss = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = L.pfann_ctx_launches(ctxh)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    losses = []
    for _ in range(args.steps):
        losses.append(step())
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    launches = L.pfann_ctx_launches(ctxh) - launches0
    prof = None
    if args.profile and rank == 0 and world == 1:
        _lib.profile(local, True)
        step()
        classes = _lib.profile_read(local)
        prof = {'classes_ms': {k: round(v[0], 3) for k, v in classes.items() if v[1]},
                'detail_ms': {k: round(v[0], 3) for k, v in _lib.profile_detail(local).items()}}
        _lib.profile(local, False)
    cpu = None
    if args.cpu_baseline and rank == 0:
        import time
        from oracle.torch_port import TorchPort, train_step_cpu   # checker / baseline only
        nrow = args.cpu_baseline
        port = TorchPort(params, {k: v.clone() for k, v in sd.items()})
        hb = {k: v[:nrow // 2].cpu() for k, v in batch.items()}
        t0 = time.perf_counter()
        lc = train_step_cpu(port, hb, tau)
        dt = time.perf_counter() - t0
        cpu = {'value': nrow / dt, 'unit': 'segments/s', 'cores': torch.get_num_threads(), 'kind': 'port',
               'sample': 'one step on %d rows of the same batch, torch-CPU fp32 autograd + Adam (%.2f s), loss %.4f' % (nrow, dt, lc)}
    if rank == 0:
        out = {
            'metric': 'training segments per second (train.py step: noise mix + impulse responses + log-mel + SpecAugment + encoder fwd/bwd '
                      '+ NT-Xent + Adam), fp32',
            'value': world * args.batch * args.steps / (ms / 1e3), 'unit': 'segments/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps,
            'scaling': 'weak', 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'train.py %s: batch %d per GPU, global NT-Xent over %d rows' %
                       (args.config, args.batch, world * args.batch), 'optimizer': 'Adam (torch)'},
            'loss_first': float(losses[0]), 'loss_last': float(losses[-1]),
            'gpu_launches': int(launches),
        }
        if cpu:
            out['cpu_baseline'] = cpu
        if prof:
            out['profile_one_step'] = prof
        if check:
            out['check'] = check
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
