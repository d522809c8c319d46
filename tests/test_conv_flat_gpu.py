"""GPU parity of the flat shifted-window tensor-core kernels (3x3 conv over zero-gapped rows, stride-2 transposed conv as
four parity classes) and the NHWC FIR + epilogue kernel, against float64 torch references on the same bf16 operands."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from brushstroke_engine_b200 import _lib
from oracle import neube_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda'
SQ2 = float(np.sqrt(2))


def md(a, b):
    return float((torch.as_tensor(a).detach().cpu().double() - torch.as_tensor(b).detach().cpu().double()).abs().max())


def pitched(x_nchw, pitch, cs, rows=None):
    """NCHW float -> zero-initialised NHWC bf16 [N, rows or H, pitch, cs] with the data in [:, :H, :W, :C]."""
    N, C, H, W = x_nchw.shape
    buf = torch.zeros((N, rows or H, pitch, cs), dtype=torch.bfloat16, device=DEV)
    buf[:, :H, :W, :C] = x_nchw.to(DEV).permute(0, 2, 3, 1).to(torch.bfloat16)
    return buf


def prep_w(w, flip):
    cout, cin = w.shape[:2]
    wd = w.to(DEV).contiguous()
    wq = torch.empty((9, cout, (cin + 63) // 64 * 64), dtype=torch.bfloat16, device=DEV)
    _lib.call('nbe_prepare_weights_bf16', _lib.ptr(wd), _lib.ptr(wq), cout, cin, 3, flip, _lib.stream())
    torch.cuda.synchronize()
    return wq


@pytest.mark.parametrize('R,cin,B,gap', [(4, 128, 3, 1), (8, 128, 5, 1), (16, 64, 2, 3), (32, 144, 2, 1), (64, 384, 1, 1), (64, 128, 3, 1), (128, 128, 1, 1)])
@pytest.mark.parametrize('valid', [0, 1])
def test_conv3x3_flat(R, cin, B, gap, valid):
    g = torch.Generator().manual_seed(R + cin + B + valid)
    IH = R + 2 if valid else R
    x = torch.randn(B, cin, IH, IH, generator=g)
    w = torch.randn(128, cin, 3, 3, generator=g) / np.sqrt(cin * 9)
    d = torch.rand(B, 128, generator=g) + 0.5
    ns = torch.rand(B, 128, generator=g) + 0.5
    noise = torch.randn(B, R, R, generator=g)
    bias = torch.randn(128, generator=g) * 0.1
    x_cs = cin + 8
    pitch = IH + gap
    xq = pitched(x, pitch, x_cs)
    wq = prep_w(w, 0)
    y_pitch = R + 2
    y = torch.full((B, R, y_pitch, 136), 5.0, dtype=torch.bfloat16, device=DEV)
    dd, nn, bb, nsd = d.to(DEV), noise.to(DEV), bias.to(DEV), ns.to(DEV)
    _lib.call('nbe_conv3x3_flat_bf16', _lib.ptr(xq), _lib.ptr(wq), _lib.ptr(y), B, R, R, cin, x_cs, pitch, valid, 128, 136,
              y_pitch, R * y_pitch, _lib.ptr(dd), _lib.ptr(nn), R * R, 0.5, _lib.ptr(bb), 0.2, SQ2, 256.0, _lib.ptr(nsd), _lib.stream())
    torch.cuda.synchronize()
    acc = F.conv2d(x.to(torch.bfloat16).double(), w.to(torch.bfloat16).double(), padding=0 if valid else 1)
    ref = acc * d.double()[:, :, None, None] + 0.5 * noise.double()[:, None] + bias.double()[None, :, None, None]
    ref = (torch.where(ref > 0, ref, ref * 0.2) * SQ2).clamp(-256, 256) * ns.double()[:, :, None, None]
    got = y[:, :, :R, :128].permute(0, 3, 1, 2).float()
    assert md(got, ref) < 1e-2 * max(float(ref.abs().max()), 1.0)
    assert float((y[:, :, R:, :].float() - 5.0).abs().max()) == 0          # gap columns of the output untouched
    assert float((y[..., 128:].float() - 5.0).abs().max()) == 0


@pytest.mark.parametrize('H,cin,B,gap', [(4, 128, 3, 1), (8, 128, 2, 2), (16, 144, 2, 1), (32, 384, 1, 1), (64, 128, 2, 1)])
def test_convT_flat_matches_conv_transpose2d(H, cin, B, gap):
    g = torch.Generator().manual_seed(H * 3 + cin)
    x = torch.randn(B, cin, H, H, generator=g)
    w = torch.randn(128, cin, 3, 3, generator=g) / np.sqrt(cin * 9)
    d = (torch.rand(B, 128, generator=g) + 0.5)
    pitch = H + gap
    xq = pitched(x, pitch, cin)
    wq = prep_w(w, 0)
    TP = 2 * H + 2
    t = torch.full((B, TP, TP, 128), 9.0, dtype=torch.bfloat16, device=DEV)
    dd = d.to(DEV)
    _lib.call('nbe_convT3x3s2_flat_bf16', _lib.ptr(xq), _lib.ptr(wq), _lib.ptr(t), B, H, H, cin, cin, pitch, 128, 128, TP, TP * TP,
              _lib.ptr(dd), _lib.stream())
    torch.cuda.synchronize()
    ref = F.conv_transpose2d(x.to(torch.bfloat16).double(), w.to(torch.bfloat16).double().transpose(0, 1), stride=2) * d.double()[:, :, None, None]
    got = t[:, :2 * H + 1, :2 * H + 1].permute(0, 3, 1, 2).float()
    assert md(got, ref) < 1e-2 * max(float(ref.abs().max()), 1.0)
    assert float((t[:, 2 * H + 1:].float() - 9.0).abs().max()) == 0 and float((t[:, :, 2 * H + 1:].float() - 9.0).abs().max()) == 0


@pytest.mark.parametrize('H,C,B', [(4, 128, 2), (16, 128, 3), (64, 128, 1)])
def test_fir_act_nhwc(H, C, B):
    g = torch.Generator().manual_seed(H + 7)
    TH = 2 * H + 1
    tt = torch.randn(B, C, TH, TH, generator=g)
    f4 = O.setup_filter([1, 3, 3, 1])
    sc = torch.rand(B, C, generator=g) + 0.5
    ns = torch.rand(B, C, generator=g) + 0.5
    noise = torch.randn(B, 2 * H, 2 * H, generator=g)
    bias = torch.randn(C, generator=g) * 0.1
    TP = TH + 1
    tq = pitched(tt, TP, C, rows=TP)
    yp = 2 * H + 1
    y = torch.full((B, 2 * H, yp, C), 3.0, dtype=torch.bfloat16, device=DEV)
    fd, scd, nsd, nd, bd = f4.to(DEV), sc.to(DEV), ns.to(DEV), noise.to(DEV), bias.to(DEV)
    _lib.call('nbe_fir_act_nhwc_bf16', _lib.ptr(tq), _lib.ptr(fd), _lib.ptr(y), B, 2 * H, 2 * H, C, TH, TH, 1, C, TP, TP * TP,
              C, yp, 2 * H * yp, 4.0, _lib.ptr(scd), _lib.ptr(nd), 4 * H * H, 0.3, _lib.ptr(bd), 0.2, SQ2, 256.0, _lib.ptr(nsd), _lib.stream())
    torch.cuda.synchronize()
    ref = O.upfirdn2d(tt.to(torch.bfloat16).double(), f4, padding=[1, 1, 1, 1], gain=4.0)
    ref = ref * sc.double()[:, :, None, None] + 0.3 * noise.double()[:, None] + bias.double()[None, :, None, None]
    ref = (torch.where(ref > 0, ref, ref * 0.2) * SQ2).clamp(-256, 256) * ns.double()[:, :, None, None]
    got = y[:, :, :2 * H].permute(0, 3, 1, 2).float()
    assert md(got, ref) < 1e-2 * max(float(ref.abs().max()), 1.0)
    assert float((y[:, :, 2 * H:].float() - 3.0).abs().max()) == 0
    # a non-separable filter takes the general 16-tap path
    fr = torch.rand(4, 4, generator=g)
    frd = fr.to(DEV)
    _lib.call('nbe_fir_act_nhwc_bf16', _lib.ptr(tq), _lib.ptr(frd), _lib.ptr(y), B, 2 * H, 2 * H, C, TH, TH, 1, C, TP, TP * TP,
              C, yp, 2 * H * yp, 1.0, None, None, 0, 0.0, None, 1.0, 1.0, -1.0, None, _lib.stream())
    torch.cuda.synchronize()
    ref = O.upfirdn2d(tt.to(torch.bfloat16).double(), fr, padding=[1, 1, 1, 1], gain=1.0)
    assert md(y[:, :, :2 * H].permute(0, 3, 1, 2).float(), ref) < 1e-2 * max(float(ref.abs().max()), 1.0)


def test_up_layer_equals_reference_formulation():
    """convT (algorithmic FLOPs) + FIR/epilogue == the reference's up-sampling modulated conv (oracle, fp64)."""
    g = torch.Generator().manual_seed(99)
    B, cin, H = 2, 144, 16
    x = torch.randn(B, cin, H, H, generator=g)
    w = torch.randn(128, cin, 3, 3, generator=g)
    s = torch.randn(B, cin, generator=g) * 0.3 + 1
    noise = torch.randn(B, 1, 2 * H, 2 * H, generator=g)
    bias = torch.randn(128, generator=g) * 0.1
    f4 = O.setup_filter([1, 3, 3, 1])
    ref = O.modulated_conv2d(x.double(), w.double(), s.double(), noise=noise.double() * 0.4, up=2, padding=1, resample_filter=f4, flip_weight=False)
    ref = O.bias_act(ref, bias.double(), act='lrelu', gain=SQ2, clamp=256)
    wsq = w.square().sum(dim=[2, 3])
    d = (s.square() @ wsq.t() + 1e-8).rsqrt()
    xq = pitched(x * s[:, :, None, None], H + 1, cin)
    wq = prep_w(w, 0)
    TP = 2 * H + 2
    t = torch.zeros((B, TP, TP, 128), dtype=torch.bfloat16, device=DEV)
    dd = d.to(DEV)
    _lib.call('nbe_convT3x3s2_flat_bf16', _lib.ptr(xq), _lib.ptr(wq), _lib.ptr(t), B, H, H, cin, cin, H + 1, 128, 128, TP, TP * TP, _lib.ptr(dd), _lib.stream())
    y = torch.zeros((B, 2 * H, 2 * H, 128), dtype=torch.bfloat16, device=DEV)
    fd, nd, bd = f4.to(DEV), noise.to(DEV).contiguous(), bias.to(DEV)
    _lib.call('nbe_fir_act_nhwc_bf16', _lib.ptr(t), _lib.ptr(fd), _lib.ptr(y), B, 2 * H, 2 * H, 128, 2 * H + 1, 2 * H + 1, 1, 128, TP, TP * TP,
              128, 2 * H, 4 * H * H, 4.0, None, _lib.ptr(nd), 4 * H * H, 0.4, _lib.ptr(bd), 0.2, SQ2, 256.0, None, _lib.stream())
    torch.cuda.synchronize()
    got = y.permute(0, 3, 1, 2).float()
    assert md(got, ref) < 2e-2 * float(ref.abs().max())


@pytest.mark.parametrize('H,cin,cout,B', [(128, 64, 128, 3), (64, 128, 256, 2), (32, 256, 256, 5), (8, 64, 128, 1), (16, 192, 128, 4)])
def test_conv3x3s2_flat_matches_strided_conv2d(H, cin, cout, B):
    """Stride-2 3x3 conv over a pre-padded NHWC input read as four parity planes == F.conv2d(stride=2) on the padded input."""
    g = torch.Generator().manual_seed(H + cin + cout + B)
    xp = torch.randn(B, cin, H + 2, H + 2, generator=g)            # already carries its 1-pixel border
    w = torch.randn(cout, cin, 3, 3, generator=g) / np.sqrt(cin * 9)
    bias = torch.randn(cout, generator=g) * 0.1
    ns = torch.rand(B, cout, generator=g) + 0.5
    xq = xp.to(DEV).permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    wq = prep_w(w, 0)
    OH = H // 2
    y_pitch, y_cs = OH + 3, cout + 8
    y = torch.full((B, OH, y_pitch, y_cs), 7.0, dtype=torch.bfloat16, device=DEV)
    bb, nsd = bias.to(DEV), ns.to(DEV)
    _lib.call('nbe_conv3x3s2_flat_bf16', _lib.ptr(xq), _lib.ptr(wq), _lib.ptr(y), B, H, H, cin, cout, y_cs, y_pitch, OH * y_pitch,
              _lib.ptr(bb), 0.01, 1.0, -1.0, _lib.ptr(nsd), _lib.stream())
    torch.cuda.synchronize()
    ref = F.conv2d(xp.to(torch.bfloat16).double(), w.to(torch.bfloat16).double(), stride=2) + bias.double()[None, :, None, None]
    ref = torch.where(ref > 0, ref, ref * 0.01) * ns.double()[:, :, None, None]
    got = y[:, :, :OH, :cout].permute(0, 3, 1, 2).float()
    assert md(got, ref) < 1e-2 * max(float(ref.abs().max()), 1.0)
    assert float((y[:, :, OH:, :].float() - 7.0).abs().max()) == 0         # columns past the output untouched
    assert float((y[..., cout:].float() - 7.0).abs().max()) == 0


def test_conv3x3_flat_two_output_passes():
    """Cout = 256 runs as two 128-channel passes; per-sample vectors are [N, Cout]."""
    g = torch.Generator().manual_seed(77)
    B, R, cin, cout = 3, 32, 64, 256
    x = torch.randn(B, cin, R + 2, R + 2, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) / np.sqrt(cin * 9)
    bias = torch.randn(cout, generator=g) * 0.1
    ns = torch.rand(B, cout, generator=g) + 0.5
    xq = pitched(x, R + 2, cin)
    wq = prep_w(w, 0)
    y = torch.full((B, R, R + 1, cout + 16), 3.0, dtype=torch.bfloat16, device=DEV)
    bb, nsd = bias.to(DEV), ns.to(DEV)
    _lib.call('nbe_conv3x3_flat_bf16', _lib.ptr(xq), _lib.ptr(wq), _lib.ptr(y), B, R, R, cin, cin, R + 2, 1, cout, cout + 16,
              R + 1, R * (R + 1), None, None, 0, 0.0, _lib.ptr(bb), 0.01, 1.0, -1.0, _lib.ptr(nsd), _lib.stream())
    torch.cuda.synchronize()
    ref = F.conv2d(x.to(torch.bfloat16).double(), w.to(torch.bfloat16).double()) + bias.double()[None, :, None, None]
    ref = torch.where(ref > 0, ref, ref * 0.01) * ns.double()[:, :, None, None]
    got = y[:, :, :R, :cout].permute(0, 3, 1, 2).float()
    assert md(got, ref) < 1e-2 * max(float(ref.abs().max()), 1.0)
    assert float((y[:, :, R:, :].float() - 3.0).abs().max()) == 0
    assert float((y[..., cout:].float() - 3.0).abs().max()) == 0


def test_random_shape_sweep():
    """Odd sizes, ragged channel counts, odd batches (dummy CTA of the last pair), all four kernels: tools/fuzz_flat.py."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, 'tools', 'fuzz_flat.py'), '40', '7'], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert 'worst relative error' in r.stdout


def test_fir_act_nhwc_cyclic_strip_order_equals_contiguous_runs():
    """At large batches the FIR pass hands out whole 16-px strips round-robin (neighbouring CTAs share halo columns through L2);
    40 images at 128^2 (320 strips >= 296 CTAs: cyclic) must give the bits of the same images filtered 20 at a time (contiguous
    tile runs) -- the arithmetic per output does not depend on the order.  Reference semantics: upfirdn2d.py:168-208."""
    B, R, C = 40, 128, 128
    g = torch.Generator().manual_seed(11)
    TP = R + 2
    T = torch.randn(B, TP, TP, C, generator=g).to(torch.bfloat16).to(DEV)
    f4 = O.setup_filter([1, 3, 3, 1]).to(DEV)
    sc = (torch.rand(B, C, generator=g) + 0.5).to(DEV)
    ns = (torch.rand(B, C, generator=g) + 0.5).to(DEV)
    noise = torch.randn(B, R, R, generator=g).to(DEV)
    bias = (torch.randn(C, generator=g) * 0.1).to(DEV)

    def run(lo, hi):
        n = hi - lo
        y = torch.zeros(n, R, R, C, dtype=torch.bfloat16, device=DEV)
        t, s, z, nz = T[lo:hi].contiguous(), sc[lo:hi].contiguous(), ns[lo:hi].contiguous(), noise[lo:hi].contiguous()
        _lib.call('nbe_fir_act_nhwc_bf16', _lib.ptr(t), _lib.ptr(f4), _lib.ptr(y), n, R, R, C, R + 1, R + 1, 1, C, TP, TP * TP,
                  C, R, R * R, 4.0, _lib.ptr(s), _lib.ptr(nz), R * R, 0.3, _lib.ptr(bias), 0.2, SQ2, 256.0, _lib.ptr(z), _lib.stream())
        torch.cuda.synchronize()
        return y

    whole = run(0, B)
    halves = torch.cat([run(0, 20), run(20, 40)])
    assert torch.equal(whole, halves)
    assert float(whole.float().abs().max()) > 0.1
