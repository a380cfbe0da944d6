#!/usr/bin/env python
"""bench.py -- the hot path of pfann on B200: fingerprints/s (builder) and queries/s vs a 10M x d128 database.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (one rank per GPU, NCCL)

Workload (BASELINE.json configs[1]): 10 000 clips x 30 s of synthetic 8 kHz mono int16
PCM per GPU -> 590 000 one-second segments -> d=128 fingerprints (configs/default.json, seeded random weights).
One "step" = one pass of mel + encoder over all clips of this rank.

  value      fingerprints/s, whole job (all ranks), PCM already resident in HBM, CUDA events, max over ranks
  e2e        same through the public host API (pfann_b200.extract.Extractor.extract_pcm16) with HOST buffers:
             H2D of the PCM from pinned memory and D2H of the fingerprints inside the timed region
  roofline   the dominant kernel (tcgen05 convolution GEMMs): algorithmic FLOPs / CUDA-event time of those
             launches, against the measured dense bf16 peak (MEASURED_PEAKS.json, sustained figure)
  match      second half of the metric: queries/s (10 s query files, 19 vectors each) vs a 10M x 128
             brute-force database row-sharded over the ranks, top-20 + sequence score, NCCL all-gather of top-k
  cpu_baseline  torch-CPU port of the same path (oracle/torch_port.py) on a bounded sample, rank 0, N=1 only

--impl reference times that CPU port alone (the reference's own code cannot travel to the GPU box; see
DESIGN.md) and prints the same line with "impl": "reference".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = 'fingerprints/s (builder) and queries/s vs 10M\u00d7d128 DB at 1/2/4/8 B200'   # BASELINE.json, both arms
FLOP_PER_SEG = 0.533e9        # algorithmic FLOPs/segment, default.json, zero-pad taps excluded
ALG_BYTES_PER_SEG = 32768 + 512   # SURVEY 8d stage 2: fp32 log-mel tile in, d=128 fp32 fingerprint out (weights amortised)
SEG_BYTES_MEL = 64768         # stage-1 algorithmic bytes per segment (fp32 entry point)
# ncu dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel family (convolutions 0-7: the fused
# layer-0 + conv kernel and the fused conv + LayerNorm kernels) for one 4096-segment chunk, per segment; see
# profiles/r02/README.md for the capture this number comes from
DOMINANT_NCU_TRAFFIC_PER_SEG = (2.0956e9 + 0.1409e9 + 4.704e9 + 2.410e9) / 4096   # front_tc (r/w) + conv_ln_tc convs 2-7 (r/w)


def workload_name(clips, clip_seconds, n_seg):
    return ('builder: %d x %d s synthetic 8 kHz mono int16 per GPU -> %d segments -> d=128 embeddings, '
            'configs/default.json, seeded random weights' % (clips, clip_seconds, n_seg))


def conv_alg_flops(params):
    """Algorithmic FLOPs per segment of each of the 16 convolutions: 2 * Ci * Co per (output
    position, kernel tap) pair whose tap reads a real input element -- taps that only see the TF-"same" zero
    padding of model.py:18-19,24-25 are not counted.  Sums to 0.533 GFLOP for configs/default.json."""
    m = params['model']
    d, h = m['d'], m['h']
    F = params['n_mels']
    T = int(params['segment_size'] * params['sample_rate']) // params['stft_hop'] + 1   # melspec.py: 1 + seg_len / hop
    ch = [1, d, d, 2 * d, 2 * d, 4 * d, 4 * d, h, h]

    def live(n):          # (output position, tap) pairs inside the input, k = 3, stride 2, TF-same padding
        no = (n - 1) // 2 + 1
        padl = ((n - 1) // 2 * 2 + 3 - n) // 2
        return sum(1 for o in range(no) for j in range(3) if 0 <= 2 * o + j - padl < n), no
    out = []
    for l in range(8):
        lt, To = live(T)
        out.append(2.0 * ch[l] * ch[l + 1] * F * lt)          # conv1: 1x3 along time
        lf, Fo = live(F)
        if m.get('fuller', False):
            out.append(2.0 * ch[l + 1] * ch[l + 1] * To * lf)  # conv2: 3x1 along frequency, dense
        else:
            out.append(2.0 * ch[l + 1] * To * lf)              # depthwise
        F, T = Fo, To
    return out


def peaks():
    p = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        j = json.load(open(p))
        return {'hbm_gbs': j['hbm_gbs'], 'tf_burst': j['bf16_tflops'], 'tf_sustained': j['bf16_tflops_sustained'],
                'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'tf_burst': 1590.0, 'tf_sustained': 1400.0, 'source': 'fallback'}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, smax, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax = max(smax, float(r[2]))
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[4:8]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': smax or None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def make_pcm_device(torch, n_clips, clip_len, seed, device):
    """Synthetic clips on the GPU: low-passed noise + 3 sinusoids in 300..3900 Hz, peak
    0.5 FS, int16.  Generated in blocks so the fp32 temporaries stay small."""
    g = torch.Generator(device=device)
    g.manual_seed(1000 + seed)
    pcm = torch.empty(n_clips * clip_len, dtype=torch.int16, device=device)
    t = torch.arange(clip_len, device=device, dtype=torch.float32) / 8000.0
    blk = 250
    for c0 in range(0, n_clips, blk):
        nb = min(blk, n_clips - c0)
        w = torch.randn((nb, clip_len + 1), generator=g, device=device)
        x = 0.6 * w[:, 1:] + 0.4 * w[:, :-1]
        for _ in range(3):
            f = 300.0 + 3600.0 * torch.rand((nb, 1), generator=g, device=device)
            a = 0.3 + 0.7 * torch.rand((nb, 1), generator=g, device=device)
            ph = 6.2831853 * torch.rand((nb, 1), generator=g, device=device)
            x = x + a * torch.sin(6.2831853 * f * t[None, :] + ph)
        x = x * (0.5 / x.abs().amax(dim=1, keepdim=True))
        pcm[c0 * clip_len:(c0 + nb) * clip_len] = torch.round(x * 32767.0).to(torch.int16).reshape(-1)
    return pcm


def timed(torch, fn, steps, warmup, barrier):
    """W untimed + K timed calls, bracketed by barrier + synchronize; device time via CUDA events."""
    for _ in range(warmup):
        fn()
    barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    barrier()
    return e0.elapsed_time(e1) / 1000.0


def cpu_port_rate(params, sd, n_seg, threads):
    """torch-CPU port (oracle/torch_port.py) on `n_seg` synthetic segments, builder batch 32 -> segments/s."""
    import torch
    from oracle.torch_port import TorchPort
    from pfann_b200 import synth
    torch.set_num_threads(threads)
    port = TorchPort(params, sd)
    rows = torch.from_numpy(synth.synth_segments(min(n_seg, 64), seed=3))
    rows = rows.repeat((n_seg + rows.shape[0] - 1) // rows.shape[0], 1)[:n_seg].contiguous()
    port.extract(rows[:64])      # warm-up (thread pools, oneDNN primitives)
    t0 = time.perf_counter()
    port.extract(rows)
    return n_seg / (time.perf_counter() - t0)


def run_reference(args):
    """--impl reference: the CPU implementation of the path on the host cores (rank 0 only)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import torch
    from pfann_b200 import synth
    params = synth.read_config('default')
    sd = synth.make_state_dict(params, seed=11)
    threads = os.cpu_count() or 1
    n_seg = args.ref_segments
    rates = []
    for i in range(args.warmup + args.steps):
        r = cpu_port_rate(params, sd, n_seg, threads)
        if i >= args.warmup:
            rates.append(r)
    v = float(np.mean(rates))
    sample = '%d segments per step (of 590000), batch 32, torch %s CPU' % (n_seg, torch.__version__)
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'fingerprints/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1000.0 * n_seg / v,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(args.clips, args.clip_seconds,
                                             args.clips * ((args.clip_seconds * 8000 - 8000) // 4000 + 1)),
                   'note': 'CPU port timed on a bounded sample of the workload'},
        'cpu_baseline': {'value': v, 'unit': 'fingerprints/s', 'cores': threads, 'kind': 'port', 'sample': sample},
        'e2e': {'value': v, 'unit': 'fingerprints/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--clips', type=int, default=10000, help='clips per GPU (config 2: 10000)')
    ap.add_argument('--clip-seconds', type=int, default=30)
    ap.add_argument('--chunk', type=int, default=16384, help='segments per internal pass')
    ap.add_argument('--precision', default='bf16')
    ap.add_argument('--db-rows', type=int, default=10_000_000)
    ap.add_argument('--queries', type=int, default=10000)
    ap.add_argument('--match-batch', type=int, default=2000, help='query files per pfann_db_query call')
    ap.add_argument('--top-k', type=int, default=20)
    ap.add_argument('--no-match', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--ref-segments', type=int, default=2048)
    ap.add_argument('--cpu-segments', type=int, default=2048)
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from pfann_b200 import _lib, synth
    from pfann_b200.extract import Extractor

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: libpfann_b200 has no CPU fallback')
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    barrier = (lambda: dist.barrier()) if world > 1 else (lambda: None)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    pk = peaks()
    params = synth.read_config('default')
    sd = synth.make_state_dict(params, seed=11)
    ex = Extractor(params, sd, device=local, precision=args.precision, chunk=args.chunk)
    clip_len = args.clip_seconds * 8000
    clip_off = np.arange(args.clips + 1, dtype=np.int64) * clip_len
    n_seg = ex.count_segments(clip_off)
    pcm = make_pcm_device(torch, args.clips, clip_len, seed=rank, device=device)
    z = torch.empty((n_seg, ex.d), dtype=torch.float32, device=device)

    # ---------------- device-resident throughput (value) + per-class kernel timing ----------------
    def step_dev():
        ex.extract_pcm16(pcm, clip_off, out=z)

    for _ in range(args.warmup):
        step_dev()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = _lib.launches(local)
    _lib.profile(local, True)
    t_dev = timed(torch, step_dev, args.steps, 0, barrier)
    prof = _lib.profile_read(local)
    detail = _lib.profile_detail(local)
    _lib.profile(local, False)
    launches = _lib.launches(local) - l0
    t_dev = max_over_ranks(t_dev)
    value = world * n_seg * args.steps / t_dev
    znorm = float(z[:1024].norm(dim=1).mean().item())

    # ---------------- end to end through the host API with host buffers (e2e) ----------------
    pcm_host = torch.empty(pcm.shape, dtype=torch.int16, pin_memory=True)
    pcm_host.copy_(pcm)
    z_host = torch.empty((n_seg, ex.d), dtype=torch.float32, pin_memory=True)
    torch.cuda.synchronize()

    def step_e2e():
        ex.extract_pcm16(pcm_host, clip_off, out=z_host)

    t_e2e = max_over_ranks(timed(torch, step_e2e, args.steps, 1, barrier))
    clocks = sampler.stop()
    e2e = world * n_seg * args.steps / t_e2e
    assert abs(float(np.linalg.norm(z_host[:16].numpy(), axis=1).mean()) - 1.0) < 1e-3
    del pcm_host

    # Dominant kernel family: the tcgen05 convolution kernels with LayerNorm + ReLU fused into the epilogue
    # (convolutions 0-7: the fused layer-0 + conv kernel and conv_ln_tc_kernel).  SURVEY 8d: stage 2 is a dense
    # contraction, bound = tensor; achieved = ALGORITHMIC FLOPs (zero-pad taps excluded) / CUDA-event time of exactly
    # those launches on the launching stream; peak = measured sustained dense bf16 (kernel timed inside a long step).
    alg = conv_alg_flops(params)
    dom = [i for i in range(8) if detail.get('conv%d' % i, (0.0, 0))[1]]
    fused_ms = sum(detail['conv%d' % i][0] for i in dom)
    fused_n = sum(detail['conv%d' % i][1] for i in dom)
    conv_ms, conv_n = prof['conv_tc']
    dom_flops = sum(alg[:8]) * n_seg * args.steps          # a fused kernel covers every convolution up to its own
    achieved = dom_flops / (fused_ms / 1000.0) / 1e12 if fused_ms > 0 else 0.0
    tens = FLOP_PER_SEG * n_seg * args.steps / (conv_ms / 1000.0) / 1e12 if conv_ms > 0 else 0.0
    outs = [524288, 262144, 131072, 65536, 65536, 32768, 16384, 8192]   # elements per segment after convolutions 0-7
    front_fused = 'conv0' not in detail                                  # layer-0 conv1 produced inside conv 1's pipeline
    act_bytes = 2 * sum(outs[i - 1] + outs[i] for i in range(2, 8)) + 2 * outs[1] + \
        (32768 if front_fused else 32768 + 4 * outs[0])                  # bf16 activations that touch HBM, in + out
    traffic = DOMINANT_NCU_TRAFFIC_PER_SEG * n_seg * args.steps / max(fused_n, 1) if DOMINANT_NCU_TRAFFIC_PER_SEG else None
    roofline = {'bound': 'tensor',
                'kernel': 'fused tcgen05 convolution + LayerNorm + ReLU kernels, convolutions 0-7 '
                          '(front_tc_kernel + conv_ln_tc_kernel)',
                'achieved': achieved, 'peak': pk['tf_sustained'], 'unit': 'TFLOP/s',
                'frac': achieved / pk['tf_sustained'], 'traffic': traffic,
                'traffic_source': 'ncu dram__bytes_read+write of these launches for one 4096-segment chunk, per segment '
                                  '(profiles/r02), scaled to the average launch of this run',
                'algorithmic_flops_per_launch': dom_flops / max(fused_n, 1),
                'algorithmic_bytes_per_segment': ALG_BYTES_PER_SEG,
                'traffic_over_algorithmic_bytes': (DOMINANT_NCU_TRAFFIC_PER_SEG / ALG_BYTES_PER_SEG
                                                   if DOMINANT_NCU_TRAFFIC_PER_SEG else None),
                'peak_source': pk['source'] + ' (bf16_tflops_sustained)',
                'launches': fused_n, 'avg_launch_ms': fused_ms / max(fused_n, 1),
                'share_of_step': fused_ms / (t_dev * 1000.0),
                # SURVEY 8d stage-2 figure over the whole step (mel, head, tail convolutions included in the time)
                'stage2_whole_step': {'tflops': value / world * FLOP_PER_SEG / 1e12,
                                      'frac_of_bf16_sustained': value / world * FLOP_PER_SEG / 1e12 / pk['tf_sustained']},
                'all_tensor_core_kernels': {'tflops': tens, 'frac_of_bf16_sustained': tens / pk['tf_sustained'],
                                            'ms_per_step': conv_ms / args.steps, 'launches': conv_n},
                # secondary view: the same launches against the HBM roofline (bf16 activations in + out, once each)
                'hbm_view': {'gbs': act_bytes * n_seg * args.steps / (fused_ms / 1000.0) / 1e9 if fused_ms > 0 else 0.0,
                             'peak_gbs': pk['hbm_gbs'], 'note': 'inter-layer activations that still touch HBM'},
                'classes_ms_per_step': {k: round(v[0] / args.steps, 3) for k, v in prof.items() if v[1]},
                'kernels_ms_per_step': {k: round(v[0] / args.steps, 3) for k, v in detail.items()}}
    mel_ms, _ = prof['mel']
    if mel_ms > 0:
        roofline['mel_hbm_frac'] = (SEG_BYTES_MEL * n_seg * args.steps / (mel_ms / 1000.0) / 1e9) / pk['hbm_gbs']
        if os.environ.get('PFANN_B200_NO_OVERLAP') is None:
            # the mel kernel of chunk k + 1 runs on a low-priority stream next to the encoder of chunk k: its event
            # pairs span the time it waits for SMs, not its work (serial run, PFANN_B200_NO_OVERLAP=1: ~52 ms per step)
            roofline['mel_note'] = ('mel runs concurrently with the encoder at low priority: its event time is elapsed '
                                    'time, not work (52 ms per 590000 segments when run alone)')

    out = {
        'metric': METRIC, 'value': value,
        'unit': 'fingerprints/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1000.0 * t_dev / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(args.clips, args.clip_seconds, n_seg),
                   'chunk': args.chunk, 'l2': 'inputs (%.1f GB PCM per GPU) larger than L2' % (pcm.numel() * 2 / 1e9),
                   'embedding_norm_check': znorm},
        'e2e': {'value': e2e, 'unit': 'fingerprints/s', 'h2d_bytes_per_step': int(pcm.numel() * 2),
                'd2h_bytes_per_step': int(n_seg * ex.d * 4), 'ms_per_step': 1000.0 * t_e2e / args.steps},
        'gpu_launches': int(launches), 'roofline': roofline, 'clocks': clocks,
    }
    del pcm, z, z_host
    torch.cuda.empty_cache()

    # ---------------- second half of the metric: queries/s vs the 10M x 128 database ----------------
    if not args.no_match:
        out['match'] = bench_match(torch, args, world, rank, local, device, barrier, max_over_ranks, pk)

    # ---------------- CPU baseline (rank 0, N = 1 only) ----------------
    if world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        v = cpu_port_rate(params, sd, args.cpu_segments, threads)
        out['cpu_baseline'] = {'value': v, 'unit': 'fingerprints/s', 'cores': threads, 'kind': 'port',
                               'sample': '%d of %d segments, builder batch 32, torch-CPU port of melspec.py+model.py '
                                         '(oracle/torch_port.py)' % (args.cpu_segments, n_seg)}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def bench_match(torch, args, world, rank, local, device, barrier, max_over_ranks, pk):
    """queries/s: 19-vector query files vs a db_rows x 128 brute-force database sharded over the ranks."""
    from pfann_b200 import _lib
    from pfann_b200.database import Database
    from pfann_b200.dist import GpuShard, ShardedDatabase, shard_songs
    d, song_len, q_len, k = 128, 59, 19, args.top_k
    n = args.db_rows
    n_songs = (n + song_len - 1) // song_len
    key = np.full(n_songs, song_len, np.int32)
    key[-1] = n - song_len * (n_songs - 1)
    pos = np.pad(np.cumsum(key, dtype=np.int64), (1, 0))
    s0, s1 = shard_songs(pos, world)[rank]
    r0, r1 = int(pos[s0]), int(pos[s1])
    g = torch.Generator(device=device)
    g.manual_seed(77 + rank)
    emb = torch.randn((r1 - r0, d), generator=g, device=device)
    emb = emb / emb.norm(dim=1, keepdim=True)
    db = Database.from_arrays(emb, key, {'top_k': k, 'frame_shift_mul': 1}, 0.5, device=local, songs=(s0, s1),
                              emb_is_shard=True)
    # queries: planted uniformly over the whole database (every shard owns its share of the true diagonals); each
    # rank contributes the rows it owns, a sum assembles them, rank 0 adds the noise and broadcasts
    nq = args.queries
    gq = torch.Generator(device=device)
    gq.manual_seed(5)
    qsong = torch.randint(0, max(1, n_songs - 1), (nq,), generator=gq, device=device)
    qoff = torch.randint(0, song_len - q_len + 1, (nq,), generator=gq, device=device)
    idx = ((qsong * song_len + qoff)[:, None] + torch.arange(q_len, device=device)[None, :]).reshape(-1)
    mine = (idx >= r0) & (idx < r1)
    q = torch.zeros((nq * q_len, d), device=device)
    q[mine] = emb[idx[mine] - r0]
    if world > 1:
        torch.distributed.all_reduce(q)
    if rank == 0:
        q = q.reshape(nq, q_len, d)
        q = q + torch.randn(q.shape, generator=gq, device=device) * (1.0 / d ** 0.5)
        q = (q / q.norm(dim=2, keepdim=True)).reshape(-1, d).contiguous()
    if world > 1:
        torch.distributed.broadcast(q, 0)
    emb_host = emb.cpu() if (world == 1 and not args.no_cpu) else None   # for the CPU baseline of this leg
    del emb
    torch.cuda.empty_cache()
    qi = np.stack([np.arange(nq, dtype=np.int64) * q_len, np.full(nq, q_len, np.int64)], axis=1)
    sdb = ShardedDatabase(GpuShard(db), k, 1, 0.5)
    B = args.match_batch
    res = {}

    def step():
        # all batches are enqueued (search, exchanges, rerank, winner combination on the device), one read-back
        _, res['song'], res['time'] = sdb.query_batches(q, qi, B)

    msteps = max(1, args.steps)
    timed(torch, step, 0, max(1, min(args.warmup, 2)), barrier)       # warm-up OUTSIDE the profiled window
    t_plain = max_over_ranks(timed(torch, step, msteps, 0, barrier))  # the reported value: no per-launch event pairs
    l0 = _lib.launches(local)
    _lib.profile(local, True)
    timed(torch, step, 1, 0, barrier)                                 # one more pass only for the per-kernel breakdown
    prof = _lib.profile_read(local)
    detail = _lib.profile_detail(local)
    _lib.profile(local, False)
    launches_per_step = int(_lib.launches(local) - l0)
    t = t_plain / msteps
    acc = float(np.mean((res['song'] == qsong.cpu().numpy()) & (res['time'] == qoff.cpu().numpy() * 0.5)))
    # per-file regime (matcher.py:136 issues one search per query file: 19 vectors per database pass): the scan is
    # HBM-bound there -- report it against the measured copy bandwidth (north_star: >= 70 % of the HBM roofline)
    q1 = q[:q_len].cpu().numpy()
    for _ in range(3):
        db.search(q1, k)
    _lib.profile(local, True)
    for _ in range(10):
        db.search(q1, k)
    _lib.profile_read(local)
    d1 = _lib.profile_detail(local)
    _lib.profile(local, False)
    pf_ms = d1.get('knn_scan_full', (0.0, 1))[0] / max(d1.get('knn_scan_full', (0.0, 1))[1], 1)
    pf_all = sum(v[0] for v in d1.values()) / 10.0
    per_file = {'queries_per_pass': q_len, 'scan_ms': pf_ms, 'search_ms_all_kernels': pf_all,
                'scan_gbs': (r1 - r0) * d * 2 / (pf_ms / 1e3) / 1e9 if pf_ms else None,
                'frac_of_hbm_peak': ((r1 - r0) * d * 2 / (pf_ms / 1e3) / 1e9) / pk['hbm_gbs'] if pf_ms else None}
    # end to end through the host API with HOST buffers (queries from pinned host memory, answers back as numpy),
    # batched regime: the same sharded call, at every N
    qh = torch.empty(q.shape, dtype=torch.float32, pin_memory=True)
    qh.copy_(q)
    torch.cuda.synchronize()
    qhn = qh.numpy()

    def step_e2e():
        sdb.query_batches(qhn, qi, B)

    step_e2e()
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step_e2e()
    torch.cuda.synchronize()
    barrier()
    e2e = nq / max_over_ranks(time.perf_counter() - t0)
    # per-file regime end to end: one call per query file, the reference's own pattern (matcher.py:136)
    nf = min(nq, 200)
    one = np.array([[0, q_len]], np.int64)
    for i in range(3):
        sdb.query_batch(qhn[i * q_len:(i + 1) * q_len], one)
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ok_pf = 0
    for i in range(nf):
        _, g1, t1 = sdb.query_batch(qhn[i * q_len:(i + 1) * q_len], one)
        ok_pf += int(g1[0] == res['song'][i] and t1[0] == res['time'][i])
    torch.cuda.synchronize()
    barrier()
    per_file['e2e_queries_per_s'] = nf / max_over_ranks(time.perf_counter() - t0)
    per_file['e2e_files'] = nf
    per_file['e2e_equals_batched'] = ok_pf == nf
    cpu = None
    if emb_host is not None:
        # reference path on the host cores: exact inner-product top-k (torch-CPU matmul + topk standing in for
        # faiss IndexFlatIP, which is not installed) + the reference's rerank arithmetic (oracle seq_score, C)
        from oracle import pfann_oracle as orc
        torch.set_num_threads(os.cpu_count() or 1)
        qh_all = q.cpu()
        dbn = emb_host.numpy()
        ncpu = min(nq, 6)
        t0 = time.perf_counter()
        ok = 0
        for i in range(ncpu):
            qq = qh_all[i * q_len:(i + 1) * q_len]
            _, lab = torch.topk(qq @ emb_host.T, k, dim=1)
            best, ss = orc.seq_score(dbn, pos, qq.numpy(), lab.numpy(), 1, 0.0)
            ok += int(best == res['song'][i] and ss[best, 1] * 0.5 == res['time'][i])
        dt = time.perf_counter() - t0
        cpu = {'value': ncpu / dt, 'unit': 'queries/s', 'cores': os.cpu_count(), 'kind': 'port',
               'sample': '%d of %d query files vs the full %d-row database; torch-CPU matmul+topk stands in for faiss '
                         'IndexFlatIP, rerank = C restatement of cpp/seqscore.cpp' % (ncpu, nq, n),
               'answers_equal_gpu': ok == ncpu}
        del emb_host, dbn
    scan_ms, scan_n = detail.get('knn_scan_full', (0.0, 0))     # filtered scans of ONE profiled step (pre-pass excluded)
    rows_local = r1 - r0
    passes = scan_n  # each launch streams the shard once against up to 256 resident queries
    scan_flops = 2.0 * nq * q_len * rows_local * d
    return {'metric': 'queries/s', 'value': nq / t, 'unit': 'queries/s',
            'e2e': {'value': e2e, 'unit': 'queries/s', 'h2d_bytes_per_step': int(nq * q_len * d * 4), 'd2h_bytes_per_step': int(nq * 12)}, 'db_rows': n, 'queries': nq,
            'vectors_per_query': q_len, 'top_k': k, 'shard_rows': rows_local, 'accuracy_vs_planted': acc,
            'query_files_per_db_pass': B, 'regime': 'batched: 256 query vectors per database pass (TMEM-read-bound epilogue, see DESIGN.md)',
            'classes_ms': {kk: round(v[0], 3) for kk, v in prof.items() if v[1]},
            'kernels_ms': {kk: [round(v[0], 3), v[1]] for kk, v in detail.items()},
            'knn_scan': {'launches': passes, 'ms': scan_ms,
                         'tflops': scan_flops / (scan_ms / 1000.0) / 1e12 if scan_ms else None,
                         'frac_of_bf16_sustained': scan_flops / (scan_ms / 1000.0) / 1e12 / pk['tf_sustained'] if scan_ms else None,
                         'gbs': passes * rows_local * d * 2 / (scan_ms / 1000.0) / 1e9 if scan_ms else None},
            'per_file_regime': per_file,
            'steps': msteps, 'ms_per_step': 1000.0 * t, 'gpu_launches': launches_per_step, 'cpu_baseline': cpu}


if __name__ == '__main__':
    main()
