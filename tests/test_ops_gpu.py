"""GPU parity: operator surface (bias_act / upfirdn2d / conv2d_resample / modulated_conv2d) through the C ABI
against the CPU oracle and the reference-generated golden fixtures."""
import numpy as np
import pytest
import torch

from conftest import load_golden, t
from oracle import neube_oracle as O
from test_oracle_golden import UPFIRDN_CASES

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def md(a, b):
    return float((torch.as_tensor(a).detach().cpu().double() - torch.as_tensor(b).detach().cpu().double()).abs().max())


def test_bias_act_golden_all_activations():
    from brushstroke_engine_b200.bias_act import bias_act
    g = load_golden('bias_act')
    x, b = t(g['x']).to(DEV), t(g['b']).to(DEV)
    for act in O.ACTIVATIONS:
        for tag, clamp in (('n', None), ('c', 0.7)):
            y = bias_act(x, b, dim=1, act=act, clamp=clamp)
            assert md(y, g[f'y_{act}_{tag}']) < 2e-6, (act, tag)
    assert md(bias_act(x, b, dim=1, act='lrelu', alpha=0.1, gain=2.5, clamp=4.0), g['y_lrelu_custom']) < 2e-6
    assert md(bias_act(t(g['x2']).to(DEV), t(g['b2']).to(DEV), dim=1, act='tanh'), g['y2_tanh']) < 2e-6


@pytest.mark.parametrize('dtype,tol', [(torch.float32, 2e-6), (torch.float16, 4e-3), (torch.bfloat16, 3e-2), (torch.float64, 2e-7)])
@pytest.mark.parametrize('shape,dim', [((3, 128, 16, 16), 1), ((2, 7, 5, 3), 1), ((5, 9), 1), ((4, 6, 10), 2), ((1, 1, 1, 1), 1), ((0, 4, 2, 2), 1)])
def test_bias_act_dtypes_shapes(dtype, tol, shape, dim):
    # float64: alpha / gain / clamp cross the ABI as float32, exactly like the reference plugin (bias_act.cpp:32)
    from brushstroke_engine_b200.bias_act import bias_act
    gen = torch.Generator().manual_seed(sum(shape) + dim)
    x = (torch.randn(shape, generator=gen) * 2).to(dtype)
    b = torch.randn(shape[dim], generator=gen).to(dtype)
    for act, kw in (('lrelu', dict(gain=np.sqrt(2), clamp=256)), ('linear', dict(clamp=0.5)), ('sigmoid', {}), ('swish', {})):
        y = bias_act(x.to(DEV), b.to(DEV), dim=dim, act=act, **kw)
        ref = O.bias_act(x.double() if dtype != torch.float32 else x, b.double() if dtype != torch.float32 else b, dim=dim, act=act, **kw)
        assert y.dtype == dtype and y.shape == x.shape
        if x.numel():
            assert md(y, ref) <= tol * max(1.0, float(ref.abs().max())), (act, dtype)


def test_bias_act_channels_last_and_noop():
    from brushstroke_engine_b200.bias_act import bias_act
    x = torch.randn(2, 8, 6, 6).to(DEV).to(memory_format=torch.channels_last)
    b = torch.randn(8).to(DEV)
    y = bias_act(x, b, act='lrelu')
    assert y.is_contiguous(memory_format=torch.channels_last)
    assert md(y, O.bias_act(x.cpu(), b.cpu(), act='lrelu')) < 2e-6
    assert bias_act(x) is x or md(bias_act(x), x) == 0          # early-out (bias_act.py:151-153)


def test_upfirdn2d_golden_cases():
    from brushstroke_engine_b200 import upfirdn2d as U
    g = load_golden('upfirdn2d')
    x = t(g['x']).to(DEV)
    for name, kw in UPFIRDN_CASES.items():
        kw = dict(kw)
        f = kw.pop('f')
        y = U.upfirdn2d(x, None if f is None else t(g[f]).to(DEV), **kw)
        assert tuple(y.shape) == g[f'y_{name}'].shape, name
        assert md(y, g[f'y_{name}']) < 3e-6, name
    f4 = t(g['f4']).to(DEV)
    assert md(U.upfirdn2d(t(g['xg']).to(DEV), f4, padding=[1, 1, 1, 1], gain=4.0), g['yg']) < 3e-6
    assert md(U.upsample2d(x, f4), g['y_upsample2d']) < 3e-6
    assert md(U.downsample2d(x, f4), g['y_downsample2d']) < 3e-6
    assert md(U.filter2d(x, f4), g['y_filter2d']) < 3e-6
    assert md(U.setup_filter([1, 3, 3, 1]), g['f4']) == 0


@pytest.mark.parametrize('dtype,tol', [(torch.float32, 3e-6), (torch.float16, 2e-3), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize('H', [4, 9, 33, 129])
def test_upfirdn2d_generator_and_upsample_shapes(dtype, tol, H):
    """(2H+1)^2 -> (2H)^2 after the transposed conv, and upsample2d x2 -- the two tiled specialisations,
    at sizes that straddle tile borders."""
    from brushstroke_engine_b200 import upfirdn2d as U
    gen = torch.Generator().manual_seed(H)
    f4 = O.setup_filter([1, 3, 3, 1])
    x = torch.randn(2, 5, H, H, generator=gen).to(dtype)
    y = U.upfirdn2d(x.to(DEV), f4.to(DEV), padding=[1, 1, 1, 1], gain=4)
    ref = O.upfirdn2d(x.float(), f4, padding=[1, 1, 1, 1], gain=4.0)
    assert y.shape == ref.shape and y.dtype == dtype
    assert md(y, ref) < tol * 4
    y = U.upsample2d(x.to(DEV), f4.to(DEV))
    ref = O.upsample2d(x.float(), f4)
    assert y.shape == ref.shape and md(y, ref) < tol * 4
    # channels-last goes through the generic strided kernel
    xc = x.to(DEV).to(memory_format=torch.channels_last)
    y = U.upfirdn2d(xc, f4.to(DEV), up=2, padding=[3, 2, 3, 2], gain=4)
    assert md(y, O.upfirdn2d(x.float(), f4, up=2, padding=[3, 2, 3, 2], gain=4.0)) < tol * 4


@pytest.mark.parametrize('dtype,tol', [(torch.float32, 3e-6), (torch.float16, 2e-3), (torch.bfloat16, 2e-2)])
def test_upfirdn2d_packed_path_padding_sweep(dtype, tol):
    """The FFMA2 path of the staged kernel (up = 1, output width a multiple of 8, paddings <= 3 incl. crops): every
    combination of left / right padding, heights that are not a multiple of the 8-row groups, several planes per CTA
    (small planes) and one strip per CTA (large planes), against the oracle."""
    from brushstroke_engine_b200 import upfirdn2d as U
    f4 = O.setup_filter([1, 3, 3, 1])
    gen = torch.Generator().manual_seed(17)
    n = 0
    for OW in (8, 16, 64, 256):
        for px0 in (-2, 0, 1, 2, 3):
            for px1 in (-1, 0, 1, 3):
                W = OW + 3 - px0 - px1
                if W < 4:
                    continue
                H = {8: 5, 16: 19, 64: 33, 256: 70}[OW]
                py0, py1 = (px1 % 3), (px0 % 4)
                x = (torch.randn(3 if OW < 256 else 1, 4 if OW < 256 else 2, H, W, generator=gen) * 3).to(dtype)
                y = U.upfirdn2d(x.to(DEV), f4.to(DEV), padding=[px0, px1, py0, py1], gain=4)
                ref = O.upfirdn2d(x.float(), f4, padding=[px0, px1, py0, py1], gain=4.0)
                assert y.shape == ref.shape and y.shape[-1] == OW
                assert md(y, ref) < tol * 12, (OW, px0, px1, H, W)
                n += 1
    assert n > 60
    # a non-finite value only reaches the outputs whose 4x4 window contains it (zero padding is a select, not a multiply)
    x = torch.randn(1, 1, 9, 9, generator=gen).to(dtype)
    x[0, 0, 3, 8] = float('inf')
    y = U.upfirdn2d(x.to(DEV), f4.to(DEV), padding=[1, 1, 1, 1], gain=4).float().cpu()
    ref = O.upfirdn2d(x.float(), f4, padding=[1, 1, 1, 1], gain=4.0)
    assert torch.equal(torch.isfinite(y), torch.isfinite(ref))


@pytest.mark.parametrize('dtype,tol', [(torch.float32, 3e-6), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize('up', [1, 2])
def test_upfirdn2d_large_and_negative_paddings(dtype, tol, up):
    """4x4 filter, dense NCHW, down = 1 with paddings outside the staged kernels' range (>= 4: e.g. filter2d(x, f4,
    padding=2) gives padx0 = 4; negative = cropping): routed to the generic kernel, checked against the oracle."""
    from brushstroke_engine_b200 import upfirdn2d as U
    f4 = O.setup_filter([1, 3, 3, 1])
    gen = torch.Generator().manual_seed(23 + up)
    x = (torch.randn(2, 3, 13, 21, generator=gen) * 2).to(dtype)
    n = 0
    for px0, px1, py0, py1 in [(4, 1, 1, 1), (4, 3, 4, 3), (5, 5, 5, 5), (8, 0, 2, 7), (0, 6, 6, 0), (3, 4, 3, 3), (2, 1, 4, 1),
                               (-3, 2, 1, 1), (1, -2, 2, 2), (2, 2, -1, 3), (1, 1, 2, -2), (-2, -1, -1, -1), (7, -1, -2, 8)]:
        y = U.upfirdn2d(x.to(DEV), f4.to(DEV), up=up, padding=[px0, px1, py0, py1], gain=up * up)
        ref = O.upfirdn2d(x.float(), f4, up=up, padding=[px0, px1, py0, py1], gain=float(up * up))
        assert y.shape == ref.shape, (px0, px1, py0, py1)
        assert md(y, ref) < tol * 12, (px0, px1, py0, py1)
        n += 1
    y = U.filter2d(x.to(DEV), f4.to(DEV), padding=2)                  # the advisor's example: padx0 = 4
    assert md(y, O.upfirdn2d(x.float(), f4, padding=[4, 3, 4, 3])) < tol * 12
    assert n == 13


def test_modconv_golden():
    from brushstroke_engine_b200.conv2d_resample import conv2d_resample
    from brushstroke_engine_b200.modconv import modulated_conv2d
    g = load_golden('modconv')
    f4 = O.setup_filter([1, 3, 3, 1]).to(DEV)
    for name, up in (('up1', 1), ('up2', 2), ('up2_odd', 2)):
        x, w, s, n = (t(g[f'{name}_{k}']).to(DEV) for k in 'xwsn')
        y = conv2d_resample(x, w, f=(f4 if up > 1 else None), up=up, padding=1, flip_weight=(up == 1))
        assert md(y, g[f'{name}_conv']) < 2e-5, name
        for demod in (True, False):
            y = modulated_conv2d(x, w, s, noise=n, up=up, padding=1, resample_filter=f4, demodulate=demod,
                                 flip_weight=(up == 1))
            assert md(y, g[f'{name}_mod_d{int(demod)}']) < 5e-5, (name, demod)


@pytest.mark.parametrize('dtype,rel', [(torch.bfloat16, 2e-2), (torch.float16, 2e-2), (torch.float32, 2e-5)])
def test_modconv_tensor_core_golden(dtype, rel):
    """The operator-surface modulated_conv2d / conv2d_resample at tensor-core shapes (3x3, Cout % 128 == 0) against the
    unmodified reference's float32 output (tests/golden/modconv_tc.npz): bf16 / fp16 activations run pack -> tcgen05 flat
    kernels (+ FIR pass for up = 2) -> unpack through nbe_modulated_conv2d, within 2e-2 of the output's max-abs (bf16
    operands, fp32 accumulation, bf16 result); float32 activations run the true-FP32 kernel through the same entry."""
    from oracle.make_golden import MODCONV_TC_CASES, modconv_tc_inputs
    from brushstroke_engine_b200.conv2d_resample import conv2d_resample
    from brushstroke_engine_b200.modconv import modulated_conv2d
    from brushstroke_engine_b200 import _lib
    g = load_golden('modconv_tc')
    f4 = O.setup_filter([1, 3, 3, 1]).to(DEV)
    for name, (N, cin, cout, H, W, up, demod, has_noise) in MODCONV_TC_CASES.items():
        x, w, s, n = modconv_tc_inputs(name)
        xd = x.to(DEV).to(dtype)
        nd = None if n is None else n.to(DEV)
        ref = t(g[f'{name}_mod'])
        scale = float(ref.abs().max())
        n0 = _lib.launch_count()
        y = modulated_conv2d(xd, w.to(DEV), s.to(DEV), noise=nd, up=up, padding=1, resample_filter=f4, demodulate=demod,
                             flip_weight=(up == 1))
        assert y.dtype == dtype and tuple(y.shape) == tuple(ref.shape), name
        assert md(y, ref) < rel * scale, (name, md(y, ref), scale)
        assert _lib.launch_count() > n0
        if f'{name}_mod_flip' in g:
            y = modulated_conv2d(xd, w.to(DEV), s.to(DEV), noise=nd, up=up, padding=1, resample_filter=f4, demodulate=demod,
                                 flip_weight=(up != 1))
            ref = t(g[f'{name}_mod_flip'])
            assert md(y, ref) < rel * float(ref.abs().max()), (name, 'flip')
            y = conv2d_resample(xd, w.to(DEV).to(dtype), f=(f4 if up > 1 else None), up=up, padding=1, flip_weight=(up == 1))
            ref = t(g[f'{name}_conv'])
            assert y.dtype == dtype and md(y, ref) < rel * float(ref.abs().max()), (name, 'conv')


def test_modconv_16bit_takes_the_tensor_core_kernels():
    """bf16 activations at a tensor-core shape must launch conv_tc_flat_kernel (not the FP32 SIMT kernel): checked through
    the entry point's own refusal of the shapes it cannot take and the workspace it asks for."""
    from brushstroke_engine_b200 import _lib
    L = _lib.load()
    assert L.nbe_modulated_conv2d_workspace(_lib.BF16, 2, 144, 8, 8, 128, 3, 1, 1) > 0
    assert L.nbe_modulated_conv2d_workspace(_lib.BF16, 2, 128, 8, 8, 3, 1, 1, 0) < 0          # ToRGB 1x1: float32 kernel
    assert L.nbe_modulated_conv2d_workspace(_lib.F32, 2, 128, 8, 8, 3, 1, 1, 0) > 0
    assert L.nbe_modulated_conv2d_workspace(_lib.BF16, 2, 128, 8, 8, 96, 3, 1, 1) < 0


@pytest.mark.parametrize('cin,cout,H,K,stride,groups', [(128, 128, 16, 3, 1, 1), (144, 128, 8, 3, 1, 1), (5, 70, 37, 3, 1, 1),
                                                         (64, 128, 16, 3, 2, 1), (1, 64, 20, 7, 1, 1), (128, 3, 16, 1, 1, 1),
                                                         (8, 12, 9, 3, 1, 4), (6, 6, 11, 5, 2, 1)])
def test_conv2d_f32_vs_torch(cin, cout, H, K, stride, groups):
    from brushstroke_engine_b200.conv2d_resample import conv2d_f32
    import torch.nn.functional as F
    gen = torch.Generator().manual_seed(cin * 1000 + cout)
    x = torch.randn(2, cin, H, H + 3, generator=gen)
    w = torch.randn(cout, cin // groups, K, K, generator=gen) / np.sqrt(cin * K * K)
    for flip in (False, True):
        y = conv2d_f32(x.to(DEV), w.to(DEV), padding=K // 2, stride=stride, groups=groups, flip=flip)
        ref = F.conv2d(x.double(), (w.flip([2, 3]) if flip else w).double(), padding=K // 2, stride=stride, groups=groups)
        assert y.shape == ref.shape
        assert md(y, ref) < 2e-5, (flip,)
