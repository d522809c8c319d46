"""``modulated_conv2d`` / ``fma`` with the reference signatures, forward only.

Mirrors thirdparty/stylegan2_ada_pytorch/training/networks.py:31-88 and torch_utils/ops/fma.py:15.
Execution: input-side modulation and epilogue demodulation (the reference's un-fused order,
networks.py:66-76, equal to the fused grouped conv to ~1e-6) folded into one ``nbe_conv2d_f32``
launch, so no per-sample weight tensor [N,O,I,k,k] is ever materialised.
"""
from __future__ import annotations

import torch

from . import _lib
from . import upfirdn2d as _up
from .conv2d_resample import conv2d_f32, conv2d_resample
from .upfirdn2d import _get_filter_size


def fma(a, b, c):
    """a * b + c (fma.py:15)."""
    return torch.addcmul(c, a, b)


def weight_sqsum(weight):
    """wsq[o, i] = sum_k w[o, i, k]^2 (float32)."""
    w = weight.to(torch.float32).contiguous()
    O, I, kh, kw = w.shape
    out = torch.empty((O, I), dtype=torch.float32, device=w.device)
    with torch.cuda.device(w.device):
        _lib.call('nbe_weight_sqsum_f32', _lib.ptr(w), _lib.ptr(out), O, I, kh * kw, _lib.stream())
    return out


def demod_coefs(styles, wsq):
    """d[n, o] = rsqrt(sum_i styles[n,i]^2 wsq[o,i] + 1e-8) (networks.py:59-64)."""
    s = styles.to(torch.float32).contiguous()
    N, I = s.shape
    O = wsq.shape[0]
    d = torch.empty((N, O), dtype=torch.float32, device=s.device)
    with torch.cuda.device(s.device):
        _lib.call('nbe_demod_coefs_f32', _lib.ptr(s), _lib.ptr(wsq), _lib.ptr(d), N, I, O, _lib.stream())
    return d


def modulated_conv2d(x, weight, styles, noise=None, up=1, down=1, padding=0, resample_filter=None, demodulate=True,
                     flip_weight=True, fused_modconv=True):
    """Same contract as the reference; ``fused_modconv`` is accepted and ignored (both of the
    reference's formulations are algebraically this one)."""
    _lib.require_cuda(x, 'modulated_conv2d')
    if (x.requires_grad or weight.requires_grad or styles.requires_grad) and torch.is_grad_enabled():
        raise RuntimeError('modulated_conv2d: forward-only op; run under torch.no_grad() '
                           '(autograd callers must keep the reference path, SURVEY.md section 3.4)')
    batch_size = x.shape[0]
    out_channels, in_channels, kh, kw = weight.shape
    assert x.ndim == 4 and x.shape[1] == in_channels
    assert styles.shape == (batch_size, in_channels)
    dtype = x.dtype
    w32 = weight.to(torch.float32).contiguous()
    s32 = styles.to(torch.float32).contiguous()
    d = demod_coefs(s32, weight_sqsum(w32)) if demodulate else None
    x32 = x.to(torch.float32)
    simple_pad = isinstance(padding, int) or len(set(padding)) == 1
    pad = padding if isinstance(padding, int) else padding[0]
    oh, ow = x.shape[2] * up // down, x.shape[3] * up // down
    nz = None if noise is None else noise.to(torch.float32)
    # noise maps [OH,OW] / [1,1,OH,OW] / [N,1,OH,OW] ride in the conv epilogue; anything else is a broadcast add
    fused_noise_ok = nz is not None and nz.numel() > 1 and nz.numel() in (oh * ow, batch_size * oh * ow) and \
        (nz.ndim < 3 or nz.shape[-3] == 1)
    if up == 1 and down == 1 and simple_pad and pad >= 0 and kh == kw:
        y = conv2d_f32(x32, w32, padding=pad, flip=not flip_weight, xscale=s32, dcoef=d,
                       noise=nz if fused_noise_ok else None)
    elif up == 2 and down == 1 and simple_pad and kh == kw and resample_filter is not None:
        fw, fh = _get_filter_size(resample_filter)
        p = [pad + (fw + up - 1) // 2, pad + (fw - up) // 2, pad + (fh + up - 1) // 2, pad + (fh - up) // 2]
        # per-channel scaling commutes with the per-channel FIR, so modulation is applied by the conv prologue
        u = _up.upfirdn2d(x32, resample_filter, up=up, padding=p, gain=up ** 2)
        y = conv2d_f32(u, w32, padding=0, flip=not flip_weight, xscale=s32, dcoef=d,
                       noise=nz if fused_noise_ok else None)
    else:
        y = conv2d_resample(x32 * s32.reshape(batch_size, -1, 1, 1), w32, f=resample_filter, up=up, down=down,
                            padding=padding, flip_weight=flip_weight)
        if d is not None:
            y = y * d.reshape(batch_size, -1, 1, 1)
        fused_noise_ok = False
    if nz is not None and not fused_noise_ok:
        y = y.add_(nz.to(y.dtype))
    return y.to(dtype)
