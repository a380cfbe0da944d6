"""Training-step pieces on the GPU (BASELINE configs[4]; reference: train.py, datautil/specaug.py,
datautil/noise.py).  Every piece is pinned to the reference's own functions run under torch autograd (tests/golden):

    y = model(x); loss.backward()               pfann_b200.model.FpNetwork in train() mode: encoder forward + backward
    loss = similarity_loss(y, tau)              train.py:41-52, forward AND backward in two CUDA launches
    loss = train_step(model, opt, x, tau, ...)  the loop body of train.py:78-104 (one GPU or one rank of many)
    x = apply_ir(x, [room, mic], pad_start)     dataset_v2.py:157-163: impulse responses, direct-form convolution
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


class ImpulseResponses:
    """Bank of impulse responses in the TIME domain, one row each, zero-padded to the longest (the reference keeps
    their spectra, datautil/ir.py:37-39,72-75; the product of spectra it forms is this bank's convolution).
    ``random_choose`` draws like ir.py:42-44 (torch's global generator, one randint call) so seeds reproduce."""

    def __init__(self, rows, device):
        L = max(len(r) for r in rows)
        bank = np.zeros((len(rows), L), np.float32)
        for i, r in enumerate(rows):
            bank[i, :len(r)] = np.asarray(r, np.float32)
        self.data = torch.from_numpy(bank).to(device)

    def random_choose(self, num):
        indices = torch.randint(0, self.data.shape[0], size=(num,), dtype=torch.long)
        return self.data[indices.to(self.data.device)]


def apply_ir(x, responses, pad_start=0, segment_size=None):
    """dataset_v2.py:157-163 for x[B, n]: convolve every row with its response(s) ``responses`` = [h[B, L], ...] one
    after the other and return samples [pad_start, segment_size) of the result (segment_size defaults to n)."""
    if not x.is_cuda:
        raise _lib.PfannError('pfann_b200.train.apply_ir needs CUDA tensors; there is no CPU fallback')
    dev = x.device.index if x.device.index is not None else torch.cuda.current_device()
    cur = x.to(torch.float32).contiguous()
    end = int(segment_size) if segment_size is not None else cur.shape[1]
    _lib.use_torch_stream(dev)
    hs = [h for h in responses if h is not None]
    for i, h in enumerate(hs):
        last = i == len(hs) - 1
        start = int(pad_start) if last else 0       # the samples before pad_start still feed the next response
        hf = h.to(torch.float32).to(x.device).contiguous()
        out = torch.empty((cur.shape[0], end - start), dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().pfann_ir_conv(_lib.ctx(dev), _lib.ptr(cur), cur.shape[0], cur.shape[1], _lib.ptr(hf),
                                            hf.shape[1], _lib.ptr(out), start, end - start), 'pfann_ir_conv')
        cur = out
    if not hs:
        cur = cur[:, int(pad_start):end].contiguous()
    return cur


def allreduce_gradients(parameters, group=None):
    """Sum the gradients over the ranks with ONE all-reduce of a flat buffer (what DistributedDataParallel's bucket
    does; the whole model is 17 M parameters at most, a single bucket).  With ``similarity_loss_gathered`` every rank
    back-propagates the GLOBAL loss through its own rows only, so the sum is exactly the gradient of the global-batch
    loss -- the same numbers one GPU computes on the whole batch, not an average of per-rank losses."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()


def train_step(model, optimizer, x, tau, specaug=None, minibatch=None, group=None):
    """One iteration of the training loop (train.py:78-104) on this rank's slice ``x`` [n][F][T] of the batch of
    log-mel segments (rows 2i, 2i+1 = a clip and its augmentation): zero_grad, SpecAugment, forward, NT-Xent over the
    global batch, backward, gradient sum over ranks, optimizer step.  ``minibatch`` < n takes the reference's
    two-pass route (train.py:83-97: fingerprints without a graph, the loss gradient, then one backward per
    minibatch).  Returns the loss as a 0-d tensor (no host sync here; train.py:105 reads it with ``.item()``)."""
    optimizer.zero_grad()
    if specaug is not None:
        x = specaug.augment(x)                                                     # train.py:81
    n = x.shape[0]
    if minibatch is not None and minibatch < n:
        with torch.no_grad():
            ys = [model(xx) for xx in torch.split(x, minibatch)]                   # train.py:85-89
        y = torch.cat(ys).requires_grad_(True)
        loss = similarity_loss_gathered(y, tau, group)
        loss.backward()
        for xx, yg in zip(torch.split(x, minibatch), torch.split(y.grad, minibatch)):
            model(xx).backward(yg)                                                 # train.py:95-97
    else:
        loss = similarity_loss_gathered(model(x), tau, group)                      # train.py:99-101
        loss.backward()
    allreduce_gradients(list(model.parameters()), group)
    optimizer.step()
    return loss.detach()
