"""Drop-in for the reference's ``model.py``: ``FpNetwork(d, h, u, F, T, params)`` with the reference's
state_dict keys, ``forward(x, norm=True)`` computed by libpfann_b200 (include/pfann_b200.h, stage 2).

    model = FpNetwork(d, h, u, F_bin, T, params['model']).to(device)              # builder.py:55
    model.load_state_dict(torch.load(os.path.join(model_dir, 'model.pt')))        # builder.py:56
    z = model(g)                                                                  # builder.py:96

The torch modules below only HOLD the parameters (so that ``load_state_dict`` / ``state_dict`` / ``.to`` /
default initialisation behave exactly like the reference's); they are never called.  ``forward`` hands the
weights to the native library once (re-done when they change) and then calls ``pfann_model_forward``.
"""
import ctypes
import os

import torch
from torch.nn import Conv1d, Conv2d, LayerNorm, Module, ModuleList

from . import _lib


def _precision(name):
    name = (name or os.environ.get('PFANN_B200_PRECISION', 'bf16')).lower()
    if name in ('bf16', 'tc', 'tensor'):
        return _lib.PRECISION_BF16
    if name in ('fp32', 'f32', 'simt'):
        return _lib.PRECISION_FP32
    raise ValueError('unknown precision %r (bf16 | fp32)' % name)


class SeparableConv2d(Module):
    """Parameter container with the shapes of model.py:15-31 (k = 3; s = (time stride of conv1, frequency stride
    of conv2))."""

    def __init__(self, i, o, k, s, in_F, in_T, fuller=False, activation='ReLU', relu_after_bn=True):
        super(SeparableConv2d, self).__init__()
        if activation not in ('ReLU', 'ELU'):
            raise KeyError(activation)                                   # model.py:12
        if k != 3:
            raise NotImplementedError('pfann_b200: kernel size 3 only (the reference never builds another, model.py:81)')
        self.conv1 = Conv2d(i, o, kernel_size=(1, k), stride=(1, s[0]))
        self.ln1 = LayerNorm((o, in_F, (in_T - 1) // s[0] + 1))
        if fuller:
            self.conv2 = Conv2d(o, o, kernel_size=(k, 1), stride=(s[1], 1))
        else:
            self.conv2 = Conv2d(o, o, kernel_size=(k, 1), stride=(s[1], 1), groups=o)
        self.ln2 = LayerNorm((o, (in_F - 1) // s[1] + 1, (in_T - 1) // s[0] + 1))


class MyF(Module):
    def __init__(self, d, h, u, in_F, in_T, fuller=False, activation='ReLU', strides=None, relu_after_bn=True):
        super(MyF, self).__init__()
        channels = [1, d, d, 2 * d, 2 * d, 4 * d, 4 * d, h, h]
        convs = []
        self.strides = []
        for i in range(8):
            s = (2, 2)
            if strides is not None:
                s = strides[i][0][1], strides[i][1][0]                   # model.py:84-85
            convs.append(SeparableConv2d(channels[i], channels[i + 1], 3, s, in_F, in_T, fuller=fuller,
                                         activation=activation, relu_after_bn=relu_after_bn))
            self.strides.append((int(s[0]), int(s[1])))
            in_F = (in_F - 1) // s[1] + 1
            in_T = (in_T - 1) // s[0] + 1
        assert in_F == in_T == 1, 'output must be 1x1'
        self.convs = ModuleList(convs)


class MyG(Module):
    def __init__(self, d, h, u):
        super(MyG, self).__init__()
        assert h % d == 0, 'h must be divisible by d'
        v = h // d
        self.d, self.h, self.u, self.v = d, h, u, v
        self.linear1 = Conv1d(d * v, d * u, kernel_size=(1,), groups=d)
        self.linear2 = Conv1d(d * u, d, kernel_size=(1,), groups=d)


class _TrainStep(torch.autograd.Function):
    """``y = model(x)`` with a backward: what torch autograd does for the reference in the training loop
    (train.py:96-103).  Forward keeps the activations inside the native model (pfann_model_train_forward), backward
    runs pfann_model_train_backward and reads one gradient per parameter in the reference's element order.  The input
    (the log-mel, nothing learnable upstream of it) gets no gradient."""

    @staticmethod
    def forward(ctx, net, x, norm, *params):
        d, h, u, F, T = net.dims
        dev = _lib.device_index(x)
        _lib.use_torch_stream(dev)          # before the handle: a parameter refresh is ordered after optimizer.step()
        hnd = net.native_handle(dev, training=True)
        xf = x.detach().reshape(-1, F, T).to(torch.float32).contiguous()
        z = torch.empty((xf.shape[0], d), dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().pfann_model_train_forward(hnd, _lib.ptr(xf), xf.shape[0], int(bool(norm)), _lib.ptr(z)),
                   'pfann_model_train_forward')
        ctx.net, ctx.dev, ctx.hnd, ctx.norm, ctx.keep = net, dev, hnd, int(bool(norm)), xf
        ctx.names = [n for n, _ in net.named_parameters()]
        ctx.shapes = [(tuple(p.shape), p.dtype) for p in params]
        return z

    @staticmethod
    def backward(ctx, dz):
        L = _lib.lib()
        _lib.use_torch_stream(ctx.dev)
        dz = dz.to(torch.float32).contiguous()
        _lib.check(L.pfann_model_train_backward(ctx.hnd, _lib.ptr(dz), ctx.norm), 'pfann_model_train_backward')
        grads = []
        for name, (shape, dtype), need in zip(ctx.names, ctx.shapes, ctx.needs_input_grad[3:]):
            if not need:
                grads.append(None)
                continue
            g = torch.empty(shape, dtype=torch.float32, device=dz.device)
            _lib.check(L.pfann_model_get_grad(ctx.hnd, name.encode(), _lib.ptr(g), g.numel()), 'pfann_model_get_grad')
            grads.append(g.to(dtype))
        return (None, None, None) + tuple(grads)


class FpNetwork(Module):
    """model.py:132-153.  ``params`` is ``params['model']`` of the JSON config; the extra optional key
    ``b200_precision`` ('bf16' tensor-core path, default; 'fp32' CUDA-core validation path) picks the kernels.
    The option variants (conv_activation='ELU', relu_after_bn=False, custom strides -- NAF-converted models) run on
    the CUDA-core kernels whatever precision is asked for: the tcgen05 path serves the default option set.
    In train() mode with gradients enabled the forward is differentiable (fp32 kernels, every option set)."""

    def __init__(self, d, h, u, F, T, params):
        super(FpNetwork, self).__init__()
        self.f = MyF(d, h, u, F, T,
                     fuller=params.get('fuller', False),
                     activation=params.get('conv_activation', 'ReLU'),
                     strides=params.get('strides'),
                     relu_after_bn=params.get('relu_after_bn', True))
        self.g = MyG(d, h, u)
        self.dims = (d, h, u, F, T)
        self.fuller = bool(params.get('fuller', False))
        self.activation = params.get('conv_activation', 'ReLU')
        self.relu_after_bn = bool(params.get('relu_after_bn', True))
        self.precision = _precision(params.get('b200_precision'))
        self.chunk = int(params.get('b200_chunk', 0))
        self._handles = {}     # device -> (handle, weights fingerprint)

    # -- native handle management ------------------------------------------------------------------
    def _fingerprint(self, training=False):
        # the training kernels are fp32: a separate native model, so that eval() between epochs (validation,
        # train.py:112-140) keeps the tensor-core one
        return tuple((p.data_ptr(), p._version) for p in self.parameters()) + (
            _lib.PRECISION_FP32 if training else self.precision, self.chunk)

    def native_handle(self, device, training=False):
        """pfann_model* for `device`, (re)built when the parameters changed since the last call."""
        fp = self._fingerprint(training)
        precision = fp[-2]
        device = (device, 'train') if training else device
        ent = self._handles.get(device)
        if ent is not None and ent[1] == fp:
            return ent[0]
        L = _lib.lib()
        if ent is None:
            d, h, u, F, T = self.dims
            hnd = ctypes.c_void_p()
            st = (ctypes.c_int * 16)(*[v for pair in self.f.strides for v in pair])
            _lib.check(L.pfann_model_create_ex(_lib.ctx(device[0] if training else device), d, h, u, F, T, int(self.fuller),
                                               {'ReLU': 0, 'ELU': 1}[self.activation], int(self.relu_after_bn), st,
                                               ctypes.byref(hnd)), 'pfann_model_create_ex')
        else:
            hnd = ent[0]
            if training and all(p.is_cuda for p in self.parameters()):
                # optimizer.step() moved the weights: refresh the kernel layouts on the device, no host round trip
                for name, p in self.named_parameters():
                    t = p.detach().to(torch.float32).contiguous()
                    _lib.check(L.pfann_model_train_load_param(hnd, name.encode(), _lib.ptr(t), t.numel()),
                               'pfann_model_train_load_param')
                self._handles[device] = (hnd, fp)
                return hnd
        for name, p in self.state_dict().items():
            t = p.detach().to(torch.float32).contiguous()
            _lib.check(L.pfann_model_set_param(hnd, name.encode(), _lib.ptr(t), t.numel()), 'pfann_model_set_param')
        if t.is_cuda:
            torch.cuda.synchronize(t.device)
        _lib.check(L.pfann_model_finalize(hnd, precision), 'pfann_model_finalize')
        if self.chunk > 0:
            _lib.check(L.pfann_model_set_chunk(hnd, self.chunk), 'pfann_model_set_chunk')
        self._handles[device] = (hnd, fp)
        return hnd

    def forward(self, x, norm=True):
        d, h, u, F, T = self.dims
        assert x.shape[-2:] == (F, T), 'expected [B, %d, %d], got %s' % (F, T, tuple(x.shape))
        if self.training and torch.is_grad_enabled():
            params = [p for _, p in self.named_parameters()]
            if any(p.requires_grad for p in params):
                return _TrainStep.apply(self, x, norm, *params)        # train.py:99-103
        dev = _lib.device_index(x)
        _lib.use_torch_stream(dev)
        # train() without a graph (the first pass of train.py:83-89): the fp32 kernels of the training model, so that
        # the fingerprints agree with what the second, differentiated pass computes; eval(): the tensor-core model
        hnd = self.native_handle(dev, training=self.training and x.is_cuda)
        xf = x.reshape(-1, F, T).to(torch.float32).contiguous()
        z = torch.empty((xf.shape[0], d), dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().pfann_model_forward(hnd, _lib.ptr(xf), xf.shape[0], int(bool(norm)), _lib.ptr(z)),
                   'pfann_model_forward')
        return z

    def layer_output(self, x, layer):
        """Parity tap: output of ``self.f.convs[layer]`` as the reference module would return it (NCHW)."""
        d, h, u, F, T = self.dims
        dev = _lib.device_index(x)
        hnd = self.native_handle(dev, training=self.training and x.is_cuda)    # the handle forward() will use
        _lib.use_torch_stream(dev)
        L = _lib.lib()
        _lib.check(L.pfann_model_set_tap(hnd, layer), 'pfann_model_set_tap')
        try:
            with torch.no_grad():
                self.forward(x)
            conv = self.f.convs[layer]
            shape = (x.shape[0],) + tuple(conv.ln2.normalized_shape)
            out = torch.empty(shape, dtype=torch.float32)
            _lib.check(L.pfann_model_get_activation(hnd, layer, _lib.ptr(out), out.numel()),
                       'pfann_model_get_activation')
        finally:
            L.pfann_model_set_tap(hnd, -1)
        return out

    def __del__(self):
        try:
            for hnd, _ in self._handles.values():
                _lib.lib().pfann_model_destroy(hnd)
        except Exception:
            pass
