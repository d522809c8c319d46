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


def _noise_layout(noise, batch_size, oh, ow):
    """-> (float32 tensor or None, per-sample stride, fused?) : maps [OH,OW] / [1,1,OH,OW] / [N,1,OH,OW] ride in the conv
    epilogue, anything else is a broadcast add afterwards."""
    if noise is None:
        return None, 0, True
    nz = noise.to(torch.float32)
    ok = nz.numel() > 1 and nz.numel() in (oh * ow, batch_size * oh * ow) and (nz.ndim < 3 or nz.shape[-3] == 1) \
        and tuple(nz.shape[-2:]) == (oh, ow)
    if not ok:
        return nz, 0, False
    return nz.contiguous(), (oh * ow if nz.numel() == batch_size * oh * ow and batch_size > 1 else 0), True


def modconv_entry(x, weight, styles, noise, up, padding, resample_filter, demodulate, flip_weight):
    """One ``nbe_modulated_conv2d`` call (include/nbe_b200.h): bf16 / fp16 activations on the tcgen05 kernels, float32 on the
    true-FP32 direct convolution.  Returns None when the combination is outside that entry point (the caller composes it
    from conv2d_resample).  ``styles`` may be None (plain conv2d_resample)."""
    lib = _lib.load()
    N, Cin, H, W = (int(v) for v in x.shape)
    Cout, K = int(weight.shape[0]), int(weight.shape[2])
    dtype = x.dtype
    code = _lib.DTYPE_CODE.get(dtype)
    if code is None or code == _lib.F64:
        return None
    f = None
    if up == 2:
        f = resample_filter
        if f is None or K != 3:
            return None
        f = f.to(torch.float32)
        if f.ndim == 1:
            f = f.ger(f)
        if tuple(f.shape) != (4, 4):
            return None
        f = f.contiguous()
    need = lib.nbe_modulated_conv2d_workspace(code, N, Cin, H, W, Cout, K, up, int(padding))
    xin = x
    if need < 0 and dtype != torch.float32:
        # 16-bit shapes the tensor-core kernels do not cover (1x1 ToRGB, Cout not a multiple of 128, ...): float32 kernel
        code = _lib.F32
        need = lib.nbe_modulated_conv2d_workspace(code, N, Cin, H, W, Cout, K, up, int(padding))
        xin = x.to(torch.float32)
    if need < 0:
        return None
    xin = xin.contiguous()
    w32 = weight.to(torch.float32).contiguous()
    s32 = None if styles is None else styles.to(torch.float32).contiguous()
    oh = H + 2 * padding - K + 1 if up == 1 else 2 * H + 2 * padding - 2
    ow = W + 2 * padding - K + 1 if up == 1 else 2 * W + 2 * padding - 2
    nz, nsn, fused = _noise_layout(noise, N, oh, ow)
    y = torch.empty((N, Cout, oh, ow), dtype=xin.dtype, device=x.device)
    wsp = torch.empty((max(int(need), 256),), dtype=torch.uint8, device=x.device)     # torch allocations are 512-byte aligned
    with torch.cuda.device(x.device):
        _lib.call('nbe_modulated_conv2d', _lib.ptr(xin), code, _lib.ptr(w32), _lib.ptr(s32), _lib.ptr(nz if fused else None), nsn,
                  _lib.ptr(y), N, Cin, H, W, Cout, K, int(up), int(padding), _lib.ptr(f), int(bool(demodulate and s32 is not None)),
                  int(bool(flip_weight)), _lib.ptr(wsp), int(wsp.numel()), _lib.stream())
    if nz is not None and not fused:
        y = y.add_(nz.to(y.dtype))
    return y.to(dtype)


@_lib.profiled('modulated_conv2d')
def modulated_conv2d(x, weight, styles, noise=None, up=1, down=1, padding=0, resample_filter=None, demodulate=True,
                     flip_weight=True, fused_modconv=True):
    """Same contract as the reference (networks.py:31-88); ``fused_modconv`` is accepted and ignored (both of the
    reference's formulations are algebraically this one).  bf16 / fp16 activations with a 3x3 kernel and a multiple of 128
    output channels run on the tensor-core kernels, float32 on the true-FP32 kernel -- see ``nbe_modulated_conv2d``."""
    _lib.require_cuda(x, 'modulated_conv2d')
    if (x.requires_grad or weight.requires_grad or styles.requires_grad) and torch.is_grad_enabled():
        raise RuntimeError('modulated_conv2d: forward-only op; run under torch.no_grad() '
                           '(autograd callers must keep the reference path, SURVEY.md section 3.4)')
    batch_size = x.shape[0]
    out_channels, in_channels, kh, kw = weight.shape
    assert x.ndim == 4 and x.shape[1] == in_channels
    assert styles.shape == (batch_size, in_channels)
    dtype = x.dtype
    simple_pad = isinstance(padding, int) or len(set(padding)) == 1
    pad = padding if isinstance(padding, int) else padding[0]
    if down == 1 and up in (1, 2) and simple_pad and pad >= 0 and kh == kw:
        y = modconv_entry(x, weight, styles, noise, up, int(pad), resample_filter, demodulate, flip_weight)
        if y is not None:
            return y
    # everything else (down-sampling, asymmetric padding, other factors): composed from conv2d_resample in float32
    w32 = weight.to(torch.float32).contiguous()
    s32 = styles.to(torch.float32).contiguous()
    d = demod_coefs(s32, weight_sqsum(w32)) if demodulate else None
    y = conv2d_resample(x.to(torch.float32) * s32.reshape(batch_size, -1, 1, 1), w32, f=resample_filter, up=up, down=down,
                        padding=padding, flip_weight=flip_weight)
    if d is not None:
        y = y * d.reshape(batch_size, -1, 1, 1)
    if noise is not None:
        y = y.add_(noise.to(y.dtype))
    return y.to(dtype)
