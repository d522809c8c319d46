"""GPU parity of the tcgen05 implicit-GEMM convolution (nbe_conv_tc_bf16) and its NHWC helpers against a float64
convolution of the same bf16-rounded operands."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from brushstroke_engine_b200 import _lib
from oracle import neube_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def md(a, b):
    return float((torch.as_tensor(a).detach().cpu().double() - torch.as_tensor(b).detach().cpu().double()).abs().max())


def pack(x_nchw, cs=None, c_off=0, scale=None):
    x_nchw = x_nchw.contiguous()
    N, C, H, W = x_nchw.shape
    cs = cs or C
    dst = torch.zeros((N, H, W, cs), dtype=torch.bfloat16, device=DEV)
    _lib.call('nbe_pack_nhwc_bf16', _lib.ptr(x_nchw), _lib.ptr(dst), N, C, H, W, cs, c_off, _lib.ptr(scale), _lib.stream())
    return dst


def unpack(x_nhwc, C):
    N, H, W, cs = x_nhwc.shape
    out = torch.empty((N, C, H, W), dtype=torch.float32, device=DEV)
    _lib.call('nbe_unpack_nchw_f32', _lib.ptr(x_nhwc), _lib.ptr(out), N, C, H, W, cs, _lib.stream())
    return out


def test_pack_unpack_roundtrip():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 37, 5, 9, generator=g).to(DEV)
    s = torch.rand(3, 37, generator=g).to(DEV)
    d = pack(x, cs=48, c_off=8, scale=s)
    ref = (x * s[:, :, None, None]).to(torch.bfloat16).float()
    assert md(d[..., 8:45].permute(0, 3, 1, 2).float(), ref) == 0
    assert float(d[..., :8].abs().max()) == 0 and float(d[..., 45:].abs().max()) == 0
    assert md(unpack(d[..., 8:].contiguous(), 37), ref) == 0


@pytest.mark.parametrize('R,cin,B', [(4, 128, 3), (8, 128, 5), (16, 128, 2), (32, 144, 2), (64, 384, 1), (128, 128, 1), (4, 128, 1), (8, 64, 17), (128, 128, 3), (128, 104, 2), (256, 128, 1)])
@pytest.mark.parametrize('valid', [False, True])
def test_conv_tc_matches_fp64_conv(R, cin, B, valid):
    g = torch.Generator().manual_seed(R * 7 + cin + B)
    cout = 128
    IH = R + 2 if valid else R
    x = torch.randn(B, cin, IH, IH, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) / np.sqrt(cin * 9)
    d = (torch.rand(B, cout, generator=g) + 0.5)
    ns = torch.rand(B, cout, generator=g) + 0.5
    noise = torch.randn(B, R, R, generator=g)
    bias = torch.randn(cout, generator=g) * 0.1
    x_cs = cin + 16                                                  # exercise a channel stride > Cin (concat buffers)
    xq = pack(x.to(DEV), cs=x_cs)
    wq = torch.empty((9, cout, (cin + 63) // 64 * 64), dtype=torch.bfloat16, device=DEV)
    wd = w.to(DEV)
    _lib.call('nbe_prepare_weights_bf16', _lib.ptr(wd), _lib.ptr(wq), cout, cin, 3, 0, _lib.stream())
    y_cs = cout + 8
    y = torch.full((B, R, R, y_cs), 7.0, dtype=torch.bfloat16, device=DEV)
    dd, nn, bb, nsd = d.to(DEV), noise.to(DEV), bias.to(DEV), ns.to(DEV)
    _lib.call('nbe_conv_tc_bf16', _lib.ptr(xq), _lib.ptr(wq), _lib.ptr(y), B, R, R, cin, x_cs, cout, y_cs, 3, int(valid),
              _lib.ptr(dd), _lib.ptr(nn), R * R, 0.5, _lib.ptr(bb), 0.2, float(np.sqrt(2)), 256.0, _lib.ptr(nsd), _lib.stream())
    torch.cuda.synchronize()
    xr = x.to(torch.bfloat16).double()
    wr = w.to(torch.bfloat16).double()
    acc = F.conv2d(xr, wr, padding=0 if valid else 1)
    ref = acc * d.double()[:, :, None, None] + 0.5 * noise.double()[:, None] + bias.double()[None, :, None, None]
    ref = torch.where(ref > 0, ref, ref * 0.2) * np.sqrt(2)
    ref = ref.clamp(-256, 256) * ns.double()[:, :, None, None]
    got = unpack(y, cout)
    scale = float(ref.abs().max())
    assert md(got, ref) < 1e-2 * max(scale, 1.0), (md(got, ref), scale)          # bf16 output rounding: 2^-9 relative
    assert float((y[..., cout:].float() - 7.0).abs().max()) == 0                 # channels beyond Cout untouched


def test_prepare_weights_flip_and_pad():
    g = torch.Generator().manual_seed(5)
    w = torch.randn(16, 20, 3, 3, generator=g)
    for flip in (0, 1):
        wq = torch.empty((9, 16, 64), dtype=torch.bfloat16, device=DEV)
        wd = w.to(DEV)
        _lib.call('nbe_prepare_weights_bf16', _lib.ptr(wd), _lib.ptr(wq), 16, 20, 3, flip, _lib.stream())
        ref = (w.flip([2, 3]) if flip else w).reshape(16, 20, 9).permute(2, 0, 1).to(torch.bfloat16).float()
        assert md(wq[:, :, :20].float(), ref) == 0
        assert float(wq[:, :, 20:].float().abs().max()) == 0


@pytest.mark.parametrize('H,C,cs', [(4, 128, 128), (16, 144, 144), (32, 384, 384), (5, 16, 24)])
def test_upsample2x_nhwc_matches_oracle(H, C, cs):
    g = torch.Generator().manual_seed(H + C)
    x = torch.randn(2, C, H, H, generator=g)
    s = torch.rand(2, C, generator=g) + 0.5
    f4 = O.setup_filter([1, 3, 3, 1])
    xq = pack(x.to(DEV), cs=cs)
    u = torch.empty((2, 2 * H + 2, 2 * H + 2, C), dtype=torch.bfloat16, device=DEV)
    fd, sd = f4.to(DEV), s.to(DEV)          # keep device copies alive across the async launch
    _lib.call('nbe_upsample2x_nhwc_bf16', _lib.ptr(xq), _lib.ptr(fd), _lib.ptr(sd), _lib.ptr(u), 2, H, H, C, cs, _lib.stream())
    ref = O.upfirdn2d(x.to(torch.bfloat16).float() * s[:, :, None, None], f4, up=2, padding=[3, 2, 3, 2], gain=4.0)
    got = u.permute(0, 3, 1, 2).float()
    assert md(got, ref) < 1e-2 * float(ref.abs().max())


def test_torgb_triad_both_layouts():
    g = torch.Generator().manual_seed(9)
    B, C, R = 2, 128, 32
    x = torch.randn(B, C, R, R, generator=g)
    w = torch.randn(3, C, generator=g)
    st = (torch.randn(B, C, generator=g) + 1) / np.sqrt(C)
    bias = torch.randn(3, generator=g) * 0.1
    colors = torch.tanh(torch.randn(B, 3, 3, generator=g))
    tt = torch.einsum('bchw,kc,bc->bkhw', x.double(), w.double(), st.double()) + bias.double()[None, :, None, None]
    uvs_ref = torch.softmax(tt.clamp(-256, 256), dim=1)
    img_ref = torch.einsum('bkhw,bck->bchw', uvs_ref, colors.double())
    wd, std, bd, cd = w.to(DEV), st.to(DEV), bias.to(DEV), colors.to(DEV)
    for is_bf16 in (0, 1):
        xin = pack(x.to(DEV)) if is_bf16 else x.to(DEV)
        img = torch.empty((B, 3, R, R), device=DEV)
        uvs = torch.empty((B, 3, R, R), device=DEV)
        _lib.call('nbe_torgb_triad', _lib.ptr(xin), is_bf16, C, _lib.ptr(wd), _lib.ptr(std), _lib.ptr(bd),
                  _lib.ptr(cd), 256.0, _lib.ptr(img), _lib.ptr(uvs), B, C, R, R, _lib.stream())
        tol = 2e-2 if is_bf16 else 1e-5
        assert md(uvs, uvs_ref) < tol and md(img, img_ref) < tol


@pytest.mark.parametrize('write_y', [0, 1])
def test_fused_torgb_epilogue_matches_separate_kernels(write_y):
    """nbe_conv_tc_bf16_torgb == nbe_conv_tc_bf16 followed by nbe_torgb_triad (ToRGB fed with the un-rounded activations)."""
    g = torch.Generator().manual_seed(21)
    B, R, C = 3, 128, 128
    x = torch.randn(B, C, R, R, generator=g)
    w = torch.randn(C, C, 3, 3, generator=g) / np.sqrt(C * 9)
    d = (torch.rand(B, C, generator=g) + 0.5).to(DEV)
    noise = torch.randn(B, R, R, generator=g).to(DEV)
    bias = (torch.randn(C, generator=g) * 0.1).to(DEV)
    rw = torch.randn(3, C, generator=g).to(DEV)
    rst = ((torch.randn(B, C, generator=g) + 1) / np.sqrt(C)).to(DEV)
    rb = (torch.randn(3, generator=g) * 0.1).to(DEV)
    col = torch.tanh(torch.randn(B, 3, 3, generator=g)).to(DEV)
    xq = pack(x.to(DEV))
    wd = w.to(DEV)
    wq = torch.empty((9, C, C), dtype=torch.bfloat16, device=DEV)
    _lib.call('nbe_prepare_weights_bf16', _lib.ptr(wd), _lib.ptr(wq), C, C, 3, 0, _lib.stream())
    sq2 = float(np.sqrt(2))
    y_ref = torch.empty((B, R, R, C), dtype=torch.bfloat16, device=DEV)
    _lib.call('nbe_conv_tc_bf16', _lib.ptr(xq), _lib.ptr(wq), _lib.ptr(y_ref), B, R, R, C, C, C, C, 3, 0, _lib.ptr(d), _lib.ptr(noise),
              R * R, 0.7, _lib.ptr(bias), 0.2, sq2, 256.0, None, _lib.stream())
    img_ref = torch.empty((B, 3, R, R), device=DEV)
    uvs_ref = torch.empty((B, 3, R, R), device=DEV)
    _lib.call('nbe_torgb_triad', _lib.ptr(y_ref), 1, C, _lib.ptr(rw), _lib.ptr(rst), _lib.ptr(rb), _lib.ptr(col), 256.0,
              _lib.ptr(img_ref), _lib.ptr(uvs_ref), B, C, R, R, _lib.stream())
    y = torch.full((B, R, R, C), 3.0, dtype=torch.bfloat16, device=DEV)
    img = torch.empty((B, 3, R, R), device=DEV)
    uvs = torch.empty((B, 3, R, R), device=DEV)
    _lib.call('nbe_conv_tc_bf16_torgb', _lib.ptr(xq), _lib.ptr(wq), _lib.ptr(y) if write_y else None, B, R, R, C, C, C, C, 0,
              _lib.ptr(d), _lib.ptr(noise), R * R, 0.7, _lib.ptr(bias), 0.2, sq2, 256.0, _lib.ptr(rw), _lib.ptr(rst), _lib.ptr(rb),
              _lib.ptr(col), 256.0, _lib.ptr(img), _lib.ptr(uvs), write_y, _lib.stream())
    torch.cuda.synchronize()
    assert md(uvs, uvs_ref) < 2e-2 and md(img, img_ref) < 2e-2        # separate path rounds the activations to bf16 first
    if write_y:
        assert md(y.float(), y_ref.float()) == 0
    else:
        assert float((y.float() - 3.0).abs().max()) == 0
    # unsupported shapes are reported, not silently mis-executed
    with pytest.raises(RuntimeError, match='fused ToRGB'):
        _lib.call('nbe_conv_tc_bf16_torgb', _lib.ptr(xq), _lib.ptr(wq), _lib.ptr(y), B * 4, 64, 64, C, C, C, C, 0,
                  _lib.ptr(d), None, 0, 0.0, _lib.ptr(bias), 0.2, sq2, 256.0, _lib.ptr(rw), _lib.ptr(rst), _lib.ptr(rb),
                  _lib.ptr(col), 256.0, _lib.ptr(img), _lib.ptr(uvs), 1, _lib.stream())
