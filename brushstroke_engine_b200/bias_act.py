"""``bias_act`` with the reference signature, executed by ``nbe_bias_act`` (sm_100a).

Mirrors thirdparty/stylegan2_ada_pytorch/torch_utils/ops/bias_act.py:55-89 (forward
only; the north star is the generator *forward* pass).  Differences by design:
no ``impl='ref'`` fallback and no silent CPU path -- non-CUDA tensors raise.
"""
from __future__ import annotations

import math

import torch

from . import _lib

# name -> (def_alpha, def_gain, cuda_idx)      bias_act.py:23-33
activation_funcs = {
    'linear':   (0.0, 1.0, 1),
    'relu':     (0.0, math.sqrt(2), 2),
    'lrelu':    (0.2, math.sqrt(2), 3),
    'tanh':     (0.0, 1.0, 4),
    'sigmoid':  (0.0, 1.0, 5),
    'elu':      (0.0, 1.0, 6),
    'selu':     (0.0, 1.0, 7),
    'softplus': (0.0, 1.0, 8),
    'swish':    (0.0, math.sqrt(2), 9),
}


@_lib.profiled('bias_act')
def bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None, impl='cuda'):
    """Fused bias + activation + gain + clamp.  Same arguments, shape, dtype and
    memory format of the result as the reference op."""
    assert isinstance(x, torch.Tensor)
    if impl != 'cuda':
        raise RuntimeError("bias_act: only impl='cuda' exists in this build (no reference fallback)")
    _lib.require_cuda(x, 'bias_act')
    if x.requires_grad and torch.is_grad_enabled():
        raise RuntimeError('bias_act: forward-only op; run under torch.no_grad() (backward is out of scope)')
    if act not in activation_funcs:
        raise KeyError(act)
    assert clamp is None or clamp >= 0
    def_alpha, def_gain, idx = activation_funcs[act]
    alpha = float(alpha if alpha is not None else def_alpha)
    gain = float(gain if gain is not None else def_gain)
    clamp = float(clamp if clamp is not None else -1)
    if x.dtype not in _lib.DTYPE_CODE:
        raise RuntimeError(f'bias_act: unsupported dtype {x.dtype}')
    # same layout handling as the reference wrapper (bias_act.py:147-149)
    channels_last = x.ndim > 2 and x.stride(1) == 1
    x = x.contiguous(memory_format=torch.channels_last if channels_last and x.ndim == 4 else torch.contiguous_format)
    size_b, step_b = 0, 1
    if b is not None:
        assert isinstance(b, torch.Tensor) and b.ndim == 1
        assert 0 <= dim < x.ndim
        assert b.shape[0] == x.shape[dim]
        if b.dtype != x.dtype or b.device != x.device:
            raise RuntimeError('bias_act: b must have the same dtype and device as x')      # bias_act.cpp:36
        b = b.contiguous()
        size_b, step_b = b.numel(), x.stride(dim)
    if act == 'linear' and gain == 1 and clamp < 0 and b is None:
        return x                                                                              # bias_act.py:151-153
    if x.numel() > 2 ** 31 - 1:
        raise RuntimeError('bias_act: x is too large')
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.call('nbe_bias_act', _lib.ptr(x), _lib.ptr(b), _lib.ptr(y), x.numel(), size_b, max(step_b, 1),
                  idx, alpha, gain, clamp, _lib.DTYPE_CODE[x.dtype], _lib.stream())
    return y
