"""``upfirdn2d`` and helpers with the reference signatures, executed by ``nbe_upfirdn2d``.

Mirrors thirdparty/stylegan2_ada_pytorch/torch_utils/ops/upfirdn2d.py:72-116 (setup_filter),
:120-164 (upfirdn2d), :272-382 (filter2d / upsample2d / downsample2d).  Forward only.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def _parse_scaling(scaling):
    if isinstance(scaling, int):
        scaling = [scaling, scaling]
    assert isinstance(scaling, (list, tuple))
    assert all(isinstance(x, int) for x in scaling)
    sx, sy = scaling
    assert sx >= 1 and sy >= 1
    return sx, sy


def _parse_padding(padding):
    if isinstance(padding, int):
        padding = [padding, padding]
    assert isinstance(padding, (list, tuple))
    assert all(isinstance(x, int) for x in padding)
    if len(padding) == 2:
        padx, pady = padding
        padding = [padx, padx, pady, pady]
    padx0, padx1, pady0, pady1 = padding
    return padx0, padx1, pady0, pady1


def _get_filter_size(f):
    if f is None:
        return 1, 1
    assert isinstance(f, torch.Tensor) and f.ndim in [1, 2]
    fw, fh = int(f.shape[-1]), int(f.shape[0])
    assert fw >= 1 and fh >= 1
    return fw, fh


def setup_filter(f, device=torch.device('cpu'), normalize=True, flip_filter=False, gain=1, separable=None):
    """upfirdn2d.py:72-116 -- float32 [fh, fw] (non-separable) or [taps] (separable)."""
    if f is None:
        f = 1
    f = torch.as_tensor(f, dtype=torch.float32)
    assert f.ndim in [0, 1, 2]
    assert f.numel() > 0
    if f.ndim == 0:
        f = f[np.newaxis]
    if separable is None:
        separable = (f.ndim == 1 and f.numel() >= 8)
    if f.ndim == 1 and not separable:
        f = f.ger(f)
    assert f.ndim == (1 if separable else 2)
    if normalize:
        f = f / f.sum()
    if flip_filter:
        f = f.flip(list(range(f.ndim)))
    f = f * (gain ** (f.ndim / 2))
    return f.to(device=device)


def _launch(x, f2d, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain):
    N, C, H, W = x.shape
    fh, fw = f2d.shape
    OW = (W * upx + padx0 + padx1 - fw + downx) // downx
    OH = (H * upy + pady0 + pady1 - fh + downy) // downy
    if OW < 1 or OH < 1:
        raise RuntimeError('upfirdn2d: output must be at least 1x1')                       # upfirdn2d.cpp:34
    channels_last = x.stride(1) == 1 and C > 1
    y = torch.empty((N, C, OH, OW), dtype=x.dtype, device=x.device,
                    memory_format=torch.channels_last if channels_last else torch.contiguous_format)
    with torch.cuda.device(x.device):
        _lib.call('nbe_upfirdn2d', _lib.ptr(x), _lib.ptr(f2d), _lib.ptr(y), N, C, H, W, *x.stride(), OH, OW, *y.stride(),
                  fh, fw, upx, upy, downx, downy, padx0, padx1, pady0, pady1, int(bool(flip)), float(gain),
                  _lib.DTYPE_CODE[x.dtype], _lib.stream())
    return y


@_lib.profiled('upfirdn2d')
def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1, impl='cuda'):
    """Pad, upsample, FIR filter and downsample a batch of 2D images (upfirdn2d.py:120-164)."""
    assert isinstance(x, torch.Tensor)
    if impl != 'cuda':
        raise RuntimeError("upfirdn2d: only impl='cuda' exists in this build (no reference fallback)")
    _lib.require_cuda(x, 'upfirdn2d')
    if x.requires_grad and torch.is_grad_enabled():
        raise RuntimeError('upfirdn2d: forward-only op; run under torch.no_grad()')
    assert x.ndim == 4
    if x.dtype not in (torch.float32, torch.float16, torch.bfloat16):
        raise RuntimeError(f'upfirdn2d: unsupported dtype {x.dtype}')
    upx, upy = _parse_scaling(up)
    downx, downy = _parse_scaling(down)
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    if f is None:
        f = torch.ones([1, 1], dtype=torch.float32, device=x.device)
    assert isinstance(f, torch.Tensor) and f.ndim in [1, 2]
    if f.dtype != torch.float32:
        raise RuntimeError('upfirdn2d: f must be float32')                                   # upfirdn2d.cpp:21
    if f.device != x.device:
        raise RuntimeError('upfirdn2d: f must reside on the same device as x')               # upfirdn2d.cpp:20
    if not (x.is_contiguous() or x.is_contiguous(memory_format=torch.channels_last)):
        x = x.contiguous()
    f = f.contiguous()
    if f.ndim == 2:
        return _launch(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip_filter, gain)
    # separable: two passes with sqrt(gain) each (upfirdn2d.py:239-240)
    y = _launch(x, f.unsqueeze(0), upx, 1, downx, 1, padx0, padx1, 0, 0, flip_filter, np.sqrt(gain))
    return _launch(y, f.unsqueeze(1), 1, upy, 1, downy, 0, 0, pady0, pady1, flip_filter, np.sqrt(gain))


def filter2d(x, f, padding=0, flip_filter=False, gain=1, impl='cuda'):
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [padx0 + fw // 2, padx1 + (fw - 1) // 2, pady0 + fh // 2, pady1 + (fh - 1) // 2]
    return upfirdn2d(x, f, padding=p, flip_filter=flip_filter, gain=gain, impl=impl)


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    upx, upy = _parse_scaling(up)
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [padx0 + (fw + upx - 1) // 2, padx1 + (fw - upx) // 2, pady0 + (fh + upy - 1) // 2, pady1 + (fh - upy) // 2]
    return upfirdn2d(x, f, up=up, padding=p, flip_filter=flip_filter, gain=gain * upx * upy, impl=impl)


def downsample2d(x, f, down=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    downx, downy = _parse_scaling(down)
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [padx0 + (fw - downx + 1) // 2, padx1 + (fw - downx) // 2, pady0 + (fh - downy + 1) // 2, pady1 + (fh - downy) // 2]
    return upfirdn2d(x, f, down=down, padding=p, flip_filter=flip_filter, gain=gain, impl=impl)
