"""Training-step pieces on the GPU (BASELINE configs[4]; reference: train.py, datautil/specaug.py,
datautil/noise.py).  The encoder backward is not built yet; what is here is the part of the step that has a clean,
reference-pinned oracle:

    loss = similarity_loss(y, tau)              train.py:41-52, forward AND backward in two CUDA launches
    x = SpecAugment(params).augment(x)          datautil/specaug.py:40-42 (one mask, like the reference)
    x = SpecAugment(params).augment_batch(x)    one mask per sample, masks drawn in the reference's RNG order
    x = add_noises(x, noise, snr_db)            datautil/noise.py:96-109 given the chosen noise rows and SNRs
"""
import numpy as np
import torch

from . import _lib


class _SimilarityLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, tau):
        if not y.is_cuda:
            raise _lib.PfannError('pfann_b200.train.similarity_loss needs a CUDA tensor; there is no CPU fallback')
        dev = y.device.index if y.device.index is not None else torch.cuda.current_device()
        yf = y.detach().to(torch.float32).contiguous()
        loss = torch.empty((), dtype=torch.float32, device=y.device)
        dy = torch.empty_like(yf)
        _lib.use_torch_stream(dev)
        _lib.check(_lib.lib().pfann_ntxent(_lib.ctx(dev), _lib.ptr(yf), yf.shape[0], yf.shape[1], float(tau),
                                           _lib.ptr(loss), _lib.ptr(dy)), 'pfann_ntxent')
        ctx.save_for_backward(dy)
        ctx.in_dtype = y.dtype
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        (dy,) = ctx.saved_tensors
        return (grad_out * dy).to(ctx.in_dtype), None


def similarity_loss(y, tau):
    """train.py:41-52: NT-Xent over y[N, d]; rows 2i and 2i+1 are a positive pair.  Differentiable."""
    return _SimilarityLoss.apply(y, tau)


def similarity_loss_gathered(y_local, tau, group=None):
    """The same loss over the GLOBAL batch when it is split across ranks (train row): all-gather the
    fingerprints, every rank evaluates the (tiny) loss on all of them and keeps the gradient rows of its own slice.
    Pairs must not straddle ranks (an even number of rows per rank)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return similarity_loss(y_local, tau)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    parts = [torch.empty_like(y_local) for _ in range(world)]
    dist.all_gather(parts, y_local.detach().contiguous(), group=group)
    parts[rank] = y_local                      # keep the graph through the local slice only
    return similarity_loss(torch.cat(parts, dim=0), tau)


class SpecAugment:
    """datautil/specaug.py:3-42.  ``get_rects`` draws exactly the random numbers ``get_mask`` draws, in the same order
    (torch's global CPU generator), and returns the three rectangles instead of a dense mask."""

    def __init__(self, params):
        self.freq_min = params.get('cutout_min', 0.1)
        self.freq_max = params.get('cutout_max', 0.5)
        self.time_min = params.get('cutout_min', 0.1)
        self.time_max = params.get('cutout_max', 0.5)
        self.cutout_min = params.get('cutout_min', 0.1)
        self.cutout_max = params.get('cutout_max', 0.5)

    def get_rects(self, F, T):
        f = int(F * (self.cutout_min + torch.rand(1) * (self.cutout_max - self.cutout_min)))      # specaug.py:18-19
        f0 = int(torch.randint(0, F - f + 1, (1,)))
        t = int(T * (self.cutout_min + torch.rand(1) * (self.cutout_max - self.cutout_min)))      # specaug.py:21-22
        t0 = int(torch.randint(0, T - t + 1, (1,)))
        fb = int(F * (self.freq_min + torch.rand(1) * (self.freq_max - self.freq_min)))           # specaug.py:27-28
        fb0 = int(torch.randint(0, F - fb + 1, (1,)))
        tb = int(T * (self.time_min + torch.rand(1) * (self.time_max - self.time_min)))           # specaug.py:33-34
        tb0 = int(torch.randint(0, T - tb + 1, (1,)))
        return [f0, f0 + f, t0, t0 + t, fb0, fb0 + fb, tb0, tb0 + tb]

    def get_mask(self, F, T):
        r = self.get_rects(F, T)
        mask = torch.zeros(F, T)
        mask[r[0]:r[1], r[2]:r[3]] = 1
        mask[r[4]:r[5], :] = 1
        mask[:, r[6]:r[7]] = 1
        return mask

    def _apply(self, x, rects):
        F, T = x.shape[-2], x.shape[-1]
        if not x.is_cuda:
            raise _lib.PfannError('pfann_b200.train.SpecAugment needs a CUDA tensor; there is no CPU fallback')
        dev = x.device.index if x.device.index is not None else torch.cuda.current_device()
        out = x.to(torch.float32).contiguous().clone()
        B = out.numel() // (F * T)
        r = torch.tensor(rects, dtype=torch.int32).reshape(-1, 8)
        if r.shape[0] == 1 and B > 1:
            r = r.expand(B, 8)
        r = r.contiguous().to(x.device)
        _lib.use_torch_stream(dev)
        _lib.check(_lib.lib().pfann_specaug_apply(_lib.ctx(dev), _lib.ptr(out), _lib.ptr(r), B, F, T),
                   'pfann_specaug_apply')
        return out

    def augment(self, x):
        """specaug.py:40-42: ONE mask for everything in x[..., F, T]."""
        return self._apply(x, [self.get_rects(x.shape[-2], x.shape[-1])])

    def augment_batch(self, x):
        """One mask per sample of x[B, ..., F, T] flattened over the leading dimensions."""
        F, T = x.shape[-2], x.shape[-1]
        B = x.numel() // (F * T)
        return self._apply(x, [self.get_rects(F, T) for _ in range(B)])


def add_noises(x, noise, snr_db):
    """datautil/noise.py:96-109 for already chosen noise rows: x[B, n] + ratio * noise[B, n] with
    ratio = rms(x) / rms(noise) * 10^(-snr / 20) per row (rms clamped at sqrt(1e-12))."""
    if not x.is_cuda:
        raise _lib.PfannError('pfann_b200.train.add_noises needs CUDA tensors; there is no CPU fallback')
    dev = x.device.index if x.device.index is not None else torch.cuda.current_device()
    xf = x.to(torch.float32).contiguous()
    nf = noise.to(torch.float32).to(x.device).contiguous()
    sf = torch.as_tensor(snr_db, dtype=torch.float32).to(x.device).contiguous()
    out = torch.empty_like(xf)
    _lib.use_torch_stream(dev)
    _lib.check(_lib.lib().pfann_snr_mix(_lib.ctx(dev), _lib.ptr(xf), _lib.ptr(nf), _lib.ptr(sf), xf.shape[0],
                                        xf.shape[1], _lib.ptr(out)), 'pfann_snr_mix')
    return out
