"""GPU: the one-kernel up-sampling layer (csrc/up_fused.cu: transposed conv on tcgen05 -> L2-resident ring of T rows -> FIR +
SynthesisLayer epilogue) against (a) the two-kernel sequence it replaces -- BIT FOR BIT, the arithmetic is the same -- and
(b) the reference's formulation in float64 (oracle: modulated_conv2d up = 2 + bias_act)."""
import numpy as np
import pytest
import torch

from brushstroke_engine_b200 import _lib
from oracle import neube_oracle as O
from test_conv_flat_gpu import md, pitched, prep_w

pytestmark = pytest.mark.gpu
DEV = 'cuda'
SQ2 = float(np.sqrt(2))


def scratch_for(W):
    n = int(_lib.load().nbe_up_layer_fused_scratch_bytes(W))
    assert n > 0
    return torch.zeros((n,), dtype=torch.uint8, device=DEV)


def two_kernels(xq, wq, fd, B, H, W, cin, y_pitch, dd, nd, nsn, ngain, bd, ns):
    TP = 2 * W + 2
    t = torch.zeros((B, 2 * H + 2, TP, 128), dtype=torch.bfloat16, device=DEV)
    _lib.call('nbe_convT3x3s2_flat_bf16', _lib.ptr(xq), _lib.ptr(wq), _lib.ptr(t), B, H, W, cin, xq.shape[3], W + 1, 128, 128, TP,
              (2 * H + 2) * TP, None, _lib.stream())
    y = torch.zeros((B, 2 * H, y_pitch, 128), dtype=torch.bfloat16, device=DEV)
    _lib.call('nbe_fir_act_nhwc_bf16', _lib.ptr(t), _lib.ptr(fd), _lib.ptr(y), B, 2 * H, 2 * W, 128, 2 * H + 1, 2 * W + 1, 1, 128, TP,
              (2 * H + 2) * TP, 128, y_pitch, 2 * H * y_pitch, 4.0, _lib.ptr(dd), _lib.ptr(nd), nsn, ngain, _lib.ptr(bd), 0.2, SQ2, 256.0,
              _lib.ptr(ns), _lib.stream())
    return y


def fused(xq, wq, fd, scratch, B, H, W, cin, y_pitch, dd, nd, nsn, ngain, bd, ns):
    y = torch.zeros((B, 2 * H, y_pitch, 128), dtype=torch.bfloat16, device=DEV)
    _lib.call('nbe_up_layer_fused_bf16', _lib.ptr(xq), _lib.ptr(wq), _lib.ptr(fd), _lib.ptr(y), _lib.ptr(scratch), scratch.numel(),
              B, H, W, cin, xq.shape[3], W + 1, 128, 128, y_pitch, 2 * H * y_pitch, 4.0, _lib.ptr(dd), _lib.ptr(nd), nsn, ngain,
              _lib.ptr(bd), 0.2, SQ2, 256.0, _lib.ptr(ns), _lib.stream())
    return y


# (B, H, W, Cin, noise: 'per' sample | 'shared' | None, output pitch slack): single-item images, runs that start in the middle of
# an image (B * items > clusters and < clusters), odd batches (a dummy image in the last pair), one and two K chunks,
# channel counts that are padded, non-square maps, the 128^2 layer's real shape
CASES = [(1, 8, 8, 128, 'per', 0), (3, 8, 8, 64, 'shared', 1), (2, 16, 16, 128, 'per', 1), (5, 16, 16, 72, None, 0),
         (40, 16, 16, 128, 'per', 1), (7, 32, 32, 128, 'per', 1), (64, 32, 32, 96, 'shared', 0), (3, 64, 64, 128, 'per', 0),
         (2, 24, 40, 128, 'per', 1), (9, 64, 64, 128, 'per', 1), (33, 40, 24, 40, 'per', 0)]


@pytest.mark.parametrize('B,H,W,cin,noise_kind,slack', CASES)
def test_fused_equals_two_kernel_sequence_bit_for_bit(B, H, W, cin, noise_kind, slack):
    g = torch.Generator().manual_seed(B * 1000 + H * 10 + cin)
    x = torch.randn(B, cin, H, W, generator=g)
    w = torch.randn(128, cin, 3, 3, generator=g) / np.sqrt(cin * 9)
    f4 = O.setup_filter([1, 3, 3, 1])
    xq = pitched(x, W + 1, (cin + 7) // 8 * 8)
    wq = prep_w(w, 0)
    fd = f4.to(DEV)
    dd = (torch.rand(B, 128, generator=g) + 0.5).to(DEV)
    bd = (torch.randn(128, generator=g) * 0.2).to(DEV)
    ns = (torch.rand(B, 128, generator=g) + 0.5).to(DEV)
    nd, nsn = None, 0
    if noise_kind == 'per':
        nd, nsn = torch.randn(B, 2 * H, 2 * W, generator=g).to(DEV), 4 * H * W
    elif noise_kind == 'shared':
        nd = torch.randn(2 * H, 2 * W, generator=g).to(DEV)
    y_pitch = 2 * W + slack
    ref = two_kernels(xq, wq, fd, B, H, W, cin, y_pitch, dd, nd, nsn, 0.37, bd, ns)
    scratch = scratch_for(W)
    for rep in range(2):                                             # the second launch runs over a used ring
        got = fused(xq, wq, fd, scratch, B, H, W, cin, y_pitch, dd, nd, nsn, 0.37, bd, ns)
        torch.cuda.synchronize()
        assert torch.equal(got, ref), (rep, float((got.float() - ref.float()).abs().max()))
    if slack:
        assert float(got[:, :, 2 * W:].abs().max()) == 0             # gap columns of the output are never written


def test_fused_up_layer_equals_reference_formulation():
    """The reference's up-sampling modulated conv + bias_act (oracle, float64) within bf16 tolerance."""
    g = torch.Generator().manual_seed(99)
    B, cin, H = 3, 128, 16
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
    y = fused(xq, wq, f4.to(DEV), scratch_for(H), B, H, H, cin, 2 * H, d.to(DEV), noise.to(DEV).contiguous(), 4 * H * H, 0.4, bias.to(DEV), None)
    torch.cuda.synchronize()
    assert md(y.permute(0, 3, 1, 2).float(), ref) < 2e-2 * float(ref.abs().max())


def test_fused_refuses_what_it_cannot_run():
    P = 4096
    for needle, args in (('Cout == 128', (P, P, P, P, P, 1 << 30, 1, 8, 8, 128, 128, 9, 256, 256, 16, 256, 4.0, None, None, 0, 0.0, None, 0.2, SQ2, 256.0, None, None)),
                         ('Cout == 128', (P, P, P, P, P, 1 << 30, 1, 8, 8, 144, 144, 9, 128, 128, 16, 256, 4.0, None, None, 0, 0.0, None, 0.2, SQ2, 256.0, None, None)),
                         ('W + 1', (P, P, P, P, P, 1 << 30, 1, 8, 8, 128, 128, 10, 128, 128, 16, 256, 4.0, None, None, 0, 0.0, None, 0.2, SQ2, 256.0, None, None)),
                         ('scratch', (P, P, P, P, P, 1024, 1, 8, 8, 128, 128, 9, 128, 128, 16, 256, 4.0, None, None, 0, 0.0, None, 0.2, SQ2, 256.0, None, None))):
        with pytest.raises(RuntimeError, match='nbe_up_layer_fused_bf16') as ei:
            _lib.call('nbe_up_layer_fused_bf16', *args)
        assert needle in str(ei.value), (needle, str(ei.value))


def test_generator_with_and_without_fused_up_layers_is_bit_identical(bundles):
    """The generator's flat path with the fused layers (b16, b128) equals the same path on the two-kernel sequence."""
    from brushstroke_engine_b200 import params as P, synthetic
    from brushstroke_engine_b200.engine import GanBrushOptions, TriadPaintEngine
    cfg, ecfg, gp, ep = bundles
    eng = TriadPaintEngine(gp, ep, DEV, mode='bf16')
    eng.G.use_up_fused = True
    B = 21
    geom = torch.from_numpy(np.concatenate([synthetic.synthetic_patch(128, seed=40 + i, radius=2 + i % 5) for i in range(B)])).to(DEV)
    opts = GanBrushOptions()
    opts.set_style(torch.cat([P.style_z_from_seed(300 + i) for i in range(B)]).to(DEV))
    opts.position = torch.from_numpy(np.random.RandomState(5).randint(0, 4000, size=(B, 2))).to(DEV)
    with torch.no_grad():
        a, ra = eng.render_tiles(geom, opts, crop_margin=10)
        a, uvs_a = a.clone(), ra['uvs'].clone()
        eng.G.use_up_fused = False
        try:
            b, rb = eng.render_tiles(geom, opts, crop_margin=10)
        finally:
            eng.G.use_up_fused = False
    assert torch.equal(a, b) and torch.equal(uvs_a, rb['uvs'])
