"""GPU parity of the block-Toeplitz form of the encoder's first layer (csrc/enc7x7_toeplitz.cu) against a float64 torch
reference on the same bf16 operands (simple_autoencoder.py:155-166: reflect-padded 7x7 conv + folded BN + LeakyReLU), of the
reflect border it writes, and against the im2col kernel it replaces."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from brushstroke_engine_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _run(x, w, bias, slope, preproc):
    B, H, W = x.shape
    lib = _lib.load()
    wt = torch.empty((7, 512, 16), dtype=torch.bfloat16, device=DEV)
    _lib.call('nbe_enc_conv7x7_toeplitz_weights', _lib.ptr(w), _lib.ptr(wt), 64, _lib.stream())
    nb = lib.nbe_enc_conv7x7_toeplitz_scratch_bytes(B, H, W)
    scratch = torch.empty(nb // 2, dtype=torch.bfloat16, device=DEV)
    y = torch.full((B, H + 2, W + 2, 64), float('nan'), dtype=torch.bfloat16, device=DEV)
    _lib.call('nbe_enc_conv7x7_toeplitz_bf16', _lib.ptr(x), _lib.ptr(wt), _lib.ptr(bias), _lib.ptr(y), _lib.ptr(scratch), nb,
              B, H, W, 64, 64, slope, preproc, _lib.stream())
    torch.cuda.synchronize()
    return y


def _reference(x, w, bias, slope, preproc):
    xx = x.double()
    if preproc == 1:
        xx = 1 - x.double()
    elif preproc == 2:
        xx = ((1 - x) * 2 - 1).double()                                 # float32 arithmetic, as base.py:32-58 on a float tensor
    xq = xx.float().to(torch.bfloat16).double()[:, None]
    wq = w.reshape(64, 1, 7, 7).to(torch.bfloat16).double()
    y = F.conv2d(F.pad(xq, (3, 3, 3, 3), mode='reflect'), wq) + bias.double()[None, :, None, None]
    return F.leaky_relu(y, slope)


@pytest.mark.parametrize('B,H,W,preproc,slope', [(3, 128, 128, 0, 0.01), (2, 16, 128, 2, 0.01), (1, 8, 256, 1, 0.2), (5, 64, 128, 0, 0.0)])
def test_toeplitz_conv7x7_matches_float64(B, H, W, preproc, slope):
    g = torch.Generator().manual_seed(B * 1000 + H + preproc)
    x = torch.rand((B, H, W), generator=g).to(DEV)
    w = (torch.randn((64, 49), generator=g) / 7).to(DEV)
    bias = (torch.randn(64, generator=g) * 0.1).to(DEV)
    y = _run(x, w, bias, slope, preproc)
    ref = _reference(x, w, bias, slope, preproc)                          # [B,64,H,W]
    got = y[:, 1:-1, 1:-1, :].permute(0, 3, 1, 2).double()
    err = (got - ref).abs()
    tol = 2.0 ** -8 * ref.abs() + 1e-5                                    # one bf16 rounding of the result
    assert bool((err <= tol).all()), float((err - tol).max())
    # the border is the reflection of the interior (padding_mode='reflect' of the next layer), bit for bit
    inner = y[:, 1:-1, 1:-1, :].permute(0, 3, 1, 2).float()
    full = F.pad(inner, (1, 1, 1, 1), mode='reflect').permute(0, 2, 3, 1).to(torch.bfloat16)
    assert torch.equal(full.view(torch.int16), y.view(torch.int16))


def test_toeplitz_conv7x7_agrees_with_im2col_kernel():
    g = torch.Generator().manual_seed(7)
    B, H, W = 4, 128, 128
    x = (torch.rand((B, H, W), generator=g) > 0.3).float().to(DEV)
    w = (torch.randn((64, 49), generator=g) / 7).to(DEV)
    bias = (torch.randn(64, generator=g) * 0.1).to(DEV)
    y = _run(x, w, bias, 0.01, 0)
    wq = torch.zeros((64, 64), dtype=torch.bfloat16, device=DEV)
    wq[:, :49] = w.to(torch.bfloat16)
    y2 = torch.zeros((B, H + 2, W + 2, 64), dtype=torch.bfloat16, device=DEV)
    _lib.call('nbe_enc_conv7x7_tc_bf16', _lib.ptr(x), _lib.ptr(wq), _lib.ptr(bias), _lib.ptr(y2), B, H, W, 64, 64, 0.01, 0, _lib.stream())
    torch.cuda.synchronize()
    a, b = y[:, 1:-1, 1:-1].float(), y2[:, 1:-1, 1:-1].float()
    # same bf16 products, another summation order: at most one bf16 ulp apart
    assert bool(((a - b).abs() <= 2.0 ** -7 * b.abs() + 1e-6).all())


def test_toeplitz_conv7x7_refuses_other_shapes():
    P = 4096
    with pytest.raises(RuntimeError, match='needs Cout == y_cs == 64'):
        _lib.call('nbe_enc_conv7x7_toeplitz_bf16', P, P, P, P, P, 1 << 30, 1, 128, 100, 64, 64, 0.01, 0, None)
    with pytest.raises(RuntimeError, match='scratch of'):
        _lib.call('nbe_enc_conv7x7_toeplitz_bf16', P, P, P, P, P, 16, 1, 128, 128, 64, 64, 0.01, 0, None)
