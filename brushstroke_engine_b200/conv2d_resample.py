"""``conv2d_resample`` with the reference signature
(thirdparty/stylegan2_ada_pytorch/torch_utils/ops/conv2d_resample.py:59-154), forward only.

The reference routes convolutions to cuDNN through ``conv2d_gradfix`` (:29-54); here they run on
``nbe_conv2d_f32`` -- a true-FP32 direct convolution (no TF32), which is what the FP32 parity
mode (<= 1e-4 against the CPU reference) needs.  Up-sampling layers use the FIR-first order
(the reference's own generic fallback, :149-154), which is algebraically identical to its
transposed-convolution fast path (:124-142) because full convolutions commute.
"""
from __future__ import annotations

import torch

from . import _lib
from . import upfirdn2d as _up
from .upfirdn2d import _get_filter_size, _parse_padding


def conv2d_f32(x, w, padding=0, stride=1, groups=1, flip=False, xscale=None, dcoef=None, noise=None, noise_gain=1.0,
               bias=None, act=0, alpha=0.0, gain=1.0, clamp=-1.0):
    """Direct FP32 convolution with the optional modulated-conv prologue / epilogue fused
    (see ``nbe_conv2d_f32`` in include/nbe_b200.h).  ``flip=False`` = correlation (F.conv2d)."""
    _lib.require_cuda(x, 'conv2d_f32')
    assert x.ndim == 4 and w.ndim == 4 and x.dtype == torch.float32 and w.dtype == torch.float32
    x = x.contiguous()
    w = w.contiguous()
    N, Cin, H, W = x.shape
    Cout, cin_g, K, K2 = w.shape
    assert K == K2, 'conv2d_f32: square kernels only'
    assert cin_g * groups == Cin, 'conv2d_f32: weight / input channel mismatch'
    OH = (H + 2 * padding - K) // stride + 1
    OW = (W + 2 * padding - K) // stride + 1
    y = torch.empty((N, Cout, OH, OW), dtype=torch.float32, device=x.device)
    noise_sn = 0
    if noise is not None:
        noise = noise.to(torch.float32).contiguous()
        if noise.numel() == OH * OW:
            noise_sn = 0
        elif noise.numel() == N * OH * OW:
            noise_sn = OH * OW
        else:
            raise RuntimeError(f'conv2d_f32: noise shape {tuple(noise.shape)} does not broadcast over channels')
    for t in (xscale, dcoef, bias):
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous() and t.device == x.device)
    with torch.cuda.device(x.device):
        _lib.call('nbe_conv2d_f32', _lib.ptr(x), _lib.ptr(w), _lib.ptr(y), N, Cin, H, W, Cout, K, int(padding), int(stride),
                  int(groups), int(bool(flip)), _lib.ptr(xscale), _lib.ptr(dcoef), _lib.ptr(noise), noise_sn, float(noise_gain),
                  _lib.ptr(bias), int(act), float(alpha), float(gain), float(clamp), _lib.stream())
    return y


def _conv(x, w, padding=0, stride=1, groups=1, flip_weight=True):
    """``_conv2d_wrapper`` (conv2d_resample.py:29-54): flip_weight=True -> correlation."""
    dtype = x.dtype
    y = conv2d_f32(x.to(torch.float32), w.to(torch.float32), padding=padding, stride=stride, groups=groups,
                   flip=not flip_weight)
    return y.to(dtype)


@_lib.profiled('conv2d_resample')
def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False):
    """2D convolution with optional up/downsampling; padding is applied once, up front."""
    assert isinstance(x, torch.Tensor) and (x.ndim == 4)
    assert isinstance(w, torch.Tensor) and (w.ndim == 4) and (w.dtype == x.dtype)
    assert f is None or (isinstance(f, torch.Tensor) and f.ndim in [1, 2] and f.dtype == torch.float32)
    assert isinstance(up, int) and (up >= 1)
    assert isinstance(down, int) and (down >= 1)
    assert isinstance(groups, int) and (groups >= 1)
    _lib.require_cuda(x, 'conv2d_resample')
    if (x.requires_grad or w.requires_grad) and torch.is_grad_enabled():
        raise RuntimeError('conv2d_resample: forward-only op; run under torch.no_grad()')
    kh, kw = int(w.shape[2]), int(w.shape[3])
    fw, fh = _get_filter_size(f)
    px0, px1, py0, py1 = _parse_padding(padding)
    # 16-bit activations, 3x3 kernel, no down-sampling: the tensor-core kernels behind nbe_modulated_conv2d (styles = NULL)
    if x.dtype in (torch.bfloat16, torch.float16) and groups == 1 and down == 1 and up in (1, 2) and kh == kw == 3 \
            and px0 == px1 == py0 == py1 and px0 >= 0 and (up == 1 or (f is not None and fw == fh == 4)):
        from .modconv import modconv_entry
        if _lib.load().nbe_modulated_conv2d_workspace(_lib.DTYPE_CODE[x.dtype], int(x.shape[0]), int(x.shape[1]), int(x.shape[2]),
                                                      int(x.shape[3]), int(w.shape[0]), 3, up, px0) >= 0:
            ff = f if (f is None or not flip_filter) else f.flip(list(range(f.ndim)))
            return modconv_entry(x, w, None, None, up, px0, ff, False, flip_weight)

    # Adjust padding to account for up/downsampling (conv2d_resample.py:94-104).
    if up > 1:
        px0 += (fw + up - 1) // 2
        px1 += (fw - up) // 2
        py0 += (fh + up - 1) // 2
        py1 += (fh - up) // 2
    if down > 1:
        px0 += (fw - down + 1) // 2
        px1 += (fw - down) // 2
        py0 += (fh - down + 1) // 2
        py1 += (fh - down) // 2

    # 1x1 convolution with downsampling only: downsample first (:107-110).
    if kw == 1 and kh == 1 and (down > 1 and up == 1):
        x = _up.upfirdn2d(x=x, f=f, down=down, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
        return _conv(x, w, groups=groups, flip_weight=flip_weight)
    # 1x1 convolution with upsampling only: convolve first (:113-116).
    if kw == 1 and kh == 1 and (up > 1 and down == 1):
        x = _conv(x, w, groups=groups, flip_weight=flip_weight)
        return _up.upfirdn2d(x=x, f=f, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    # Downsampling only: FIR then strided convolution (:119-122).
    if down > 1 and up == 1:
        x = _up.upfirdn2d(x=x, f=f, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
        return _conv(x, w, stride=down, groups=groups, flip_weight=flip_weight)
    # No resampling and conv-expressible padding (:145-147).
    if up == 1 and down == 1 and px0 == px1 == py0 == py1 and px0 >= 0:
        return _conv(x, w, padding=px0, groups=groups, flip_weight=flip_weight)
    # Everything else, including the generator's up=2 layers: FIR-upsample, convolve, FIR-downsample (:149-154).
    x = _up.upfirdn2d(x=x, f=(f if up > 1 else None), up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    x = _conv(x, w, groups=groups, flip_weight=flip_weight)
    if down > 1:
        x = _up.upfirdn2d(x=x, f=f, down=down, flip_filter=flip_filter)
    return x
