"""CPU: the oracle restatement against fixtures produced by the unmodified reference
(oracle/make_golden.py).  No GPU, no /root/reference."""
import numpy as np
import pytest
import torch

from conftest import load_golden, t
from oracle import neube_oracle as O
from brushstroke_engine_b200 import params as P


def md(a, b):
    return float((torch.as_tensor(a).double() - torch.as_tensor(b).double()).abs().max())


def test_bias_act_all_activations():
    g = load_golden('bias_act')
    x, b = t(g['x']), t(g['b'])
    for act in O.ACTIVATIONS:
        for tag, clamp in (('n', None), ('c', 0.7)):
            assert md(O.bias_act(x, b, dim=1, act=act, clamp=clamp), g[f'y_{act}_{tag}']) < 1e-6, (act, tag)
    assert md(O.bias_act(x, b, dim=1, act='lrelu', alpha=0.1, gain=2.5, clamp=4.0), g['y_lrelu_custom']) < 1e-6
    assert md(O.bias_act(t(g['x2']), t(g['b2']), dim=1, act='tanh'), g['y2_tanh']) < 1e-6


UPFIRDN_CASES = {
    'gen':   dict(f='f4', up=1, down=1, padding=[1, 1, 1, 1], flip_filter=False, gain=4.0),
    'up2':   dict(f='f4', up=2, down=1, padding=[2, 1, 2, 1], flip_filter=False, gain=4.0),
    'down2': dict(f='f4', up=1, down=2, padding=[1, 1, 1, 1], flip_filter=False, gain=1.0),
    'asym':  dict(f='fa', up=[3, 2], down=[2, 1], padding=[2, -1, 0, 3], flip_filter=True, gain=1.5),
    'neg':   dict(f='fa', up=1, down=1, padding=[-1, 2, -2, 3], flip_filter=False, gain=1.0),
    'sep':   dict(f='f1', up=2, down=1, padding=[4, 3, 4, 3], flip_filter=False, gain=4.0),
    'none':  dict(f=None, up=2, down=1, padding=0, flip_filter=False, gain=1.0),
}


def test_upfirdn2d_cases():
    g = load_golden('upfirdn2d')
    x = t(g['x'])
    assert md(O.setup_filter([1, 3, 3, 1]), g['f4']) == 0
    for name, kw in UPFIRDN_CASES.items():
        kw = dict(kw)
        f = kw.pop('f')
        y = O.upfirdn2d(x, None if f is None else t(g[f]), **kw)
        assert y.shape == g[f'y_{name}'].shape, name
        assert md(y, g[f'y_{name}']) < 2e-6, name
    f4 = t(g['f4'])
    assert md(O.upfirdn2d(t(g['xg']), f4, padding=[1, 1, 1, 1], gain=4.0), g['yg']) < 2e-6
    assert md(O.upsample2d(x, f4), g['y_upsample2d']) < 2e-6
    assert md(O.downsample2d(x, f4), g['y_downsample2d']) < 2e-6
    assert md(O.filter2d(x, f4), g['y_filter2d']) < 2e-6


def test_modulated_conv2d():
    g = load_golden('modconv')
    f4 = O.setup_filter([1, 3, 3, 1])
    for name, up in (('up1', 1), ('up2', 2), ('up2_odd', 2)):
        x, w, s, n = (t(g[f'{name}_{k}']) for k in 'xwsn')
        y = O.conv2d_resample(x, w, f=(f4 if up > 1 else None), up=up, padding=1, flip_weight=(up == 1))
        assert md(y, g[f'{name}_conv']) < 2e-5, name
        for demod in (True, False):
            y = O.modulated_conv2d(x, w, s, noise=n, up=up, padding=1, resample_filter=f4, demodulate=demod,
                                   flip_weight=(up == 1))
            assert md(y, g[f'{name}_mod_d{int(demod)}']) < 5e-5, (name, demod)


def test_generator_and_encoder(bundles):
    cfg, ecfg, gp, ep = bundles
    g = load_golden('generator')
    z, geom, pos = t(g['z']), t(g['geom']), t(g['positions'])
    gf = O.geometry_encode(ep, ecfg, geom)
    assert md(gf[0], g['g0']) < 2e-5
    assert md(gf[1][:, ::8], g['g1_sub']) < 2e-5
    for tag, positions in (('nopos', None), ('pos', pos)):
        img, dbg = O.generator_forward(gp, cfg, z, gf, positions=positions, return_features=[64])
        assert md(dbg['ws'], g['ws']) < 1e-5
        assert md(img, g[f'img32_{tag}']) < 1e-4
        assert md(dbg['uvs'], g[f'uvs32_{tag}']) < 1e-4
        assert md(dbg['colors'], g[f'colors32_{tag}']) < 1e-5
        assert md(dbg['features64'][:, ::16, ::2, ::2], g[f'feat64_sub_{tag}']) < 1e-4
        # the reference's own mixed-fp16 output stays within the BF16-mode tolerance of the fp32 oracle
        assert md(img, g[f'img16_{tag}'].astype(np.float32)) < 2e-2
    npos = (pos % 128) / 127
    nc = gp['synthesis.b16.conv1.noise_const']
    assert md(O.shifted_noise(nc, npos), g['noise16_pos']) < 1e-6
    assert md(O.shifted_noise_closed_form(nc, npos), g['noise16_pos']) < 1e-5


def test_encoder_neg_slope_variant():
    """--neg_slope autoencoder (conv -> LeakyReLU -> BatchNorm, ScaleUpV2; simple_autoencoder.py:48-53,128-148): the oracle's
    restatement against features produced by the reference's own factory-built model (oracle/make_golden.py --only encoder_v2)."""
    g = load_golden('encoder_v2')
    ecfg = P.EncoderConfig(bn_after_activation=True, neg_slope=0.2)
    ep = P.init_encoder_params(ecfg, seed=5, perturb_bn=0.1)
    assert bytes(g['enc_digest']).decode() == P.bundle_digest(ep)
    gf = O.geometry_encode(ep, ecfg, t(g['geom']))
    assert md(gf[0], g['g0']) < 2e-5
    assert md(gf[1][:, ::8], g['g1_sub']) < 2e-5


def test_stylizer_engine(bundles):
    cfg, ecfg, gp, ep = bundles
    g = load_golden('engine')
    guidance = g['guidance']
    z = P.style_z_from_seed(594)
    for level, mode in ((0, 'clear'), (0, 'full'), (2, 'clear')):
        canvas, crops, metas = O.stylize(gp, ep, cfg, ecfg, guidance, z, 10, 'all', mode, level)
        assert np.array_equal(np.array(crops, dtype=np.int32), g['crops'])          # bit-exact patch indexing
        assert np.array_equal(np.array(metas, dtype=np.int32), g['metas'])          # bit-exact tile placement
        diff = np.abs(canvas.astype(np.int32) - g[f'canvas_l{level}_{mode}'].astype(np.int32))
        assert diff.max() <= 1 and (diff > 0).mean() < 1e-3, (level, mode, diff.max())
    sf = O.uvs_sfactor  # noqa: F841  (exercised through map_style_s below)
    mapped = O.map_style_s(torch.tensor(float(g['sfactor'])), t(g['uvs5_sub']))
    assert md(mapped, g['mapped5_sub']) < 1e-6


def test_crop_grid_sizes():
    """2000^2 -> 529 crops / 2152^2 canvas; 4096^2 -> 2209 crops / 4264^2."""
    for size, ncrops, canvas in ((2000, 529, 2152), (4096, 2209, 4264)):
        geo = O.pad_geo(np.full((size, size, 1), 255, np.uint8), 10)
        crops, padded = O.generate_stitching_crops(geo, 128, 'all', 20)
        assert len(crops) == ncrops and padded.shape[:2] == (canvas, canvas)


def test_two_restatements_agree():
    """conv-based vs tap-by-tap upfirdn2d; transposed-conv vs FIR-first conv2d_resample (what the CUDA path runs)."""
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 3, 9, 11, generator=g)
    fa = torch.rand(3, 5, generator=g)
    f4 = O.setup_filter([1, 3, 3, 1])
    for kw in (dict(up=[3, 2], down=[2, 1], padding=[2, -1, 0, 3], flip_filter=True, gain=1.5),
               dict(up=2, padding=[3, 2, 3, 2], gain=4.0), dict(down=2, padding=[1, 1, 1, 1])):
        f = fa if 'flip_filter' in kw else f4
        assert md(O.upfirdn2d(x, f, **kw), O.upfirdn2d_closed_form(x, f, **kw)) < 2e-6
    w = torch.randn(5, 3, 3, 3, generator=g)
    for up, fl in ((2, False), (2, True), (1, True), (1, False)):
        a = O.conv2d_resample(x, w, f=f4 if up > 1 else None, up=up, padding=1, flip_weight=fl)
        b = O.conv2d_resample_definition(x, w, f=f4 if up > 1 else None, up=up, padding=1, flip_weight=fl)
        assert a.shape == b.shape and md(a, b) < 2e-5


def test_canvas_colour_format(bundles):
    """'canvas' colour format: 3 + 5 ToRGB outputs, canvas/alpha blend (networks.py:433-481) and the four render modes of
    ``CanvasPaintEngine`` (brush.py:870-935) against the reference-generated fixture."""
    _, ecfg, _, ep = bundles
    g = load_golden('canvas')
    cfg = P.GeneratorConfig(color_format='canvas')
    gp = P.init_generator_params(cfg, seed=3, perturb=0.1)
    assert P.bundle_digest(gp) == bytes(g['gen_digest']).decode()
    z, geom = t(g['z']), t(g['geom'])
    gf = O.geometry_encode(ep, ecfg, geom)
    img, d = O.generator_forward(gp, cfg, z, gf)
    assert md(img, g['img32']) < 1e-4
    assert md(d['uvs'][:, :, ::2, ::2], g['uvs32_sub']) < 1e-4 and md(d['colors'], g['colors32']) < 1e-5
    assert md(d['canvas'][:, :, ::2, ::2], g['canvas32_sub']) < 1e-4 and md(d['alpha'][:, :, ::2, ::2], g['alpha32_sub']) < 1e-4
    for mode in ('clear', 'stroke', 'canvas', 'full'):
        rgba = O.canvas_composite(d['uvs'], d['colors'], d['alpha_fg'], d['canvas'], mode, color1=torch.tensor([1.0, 0.0, 128 / 255]))
        assert md(rgba[:, :, ::2, ::2], g[f'rgba_{mode}_sub']) < 1e-4, mode
    with pytest.raises(RuntimeError):
        P.GeneratorConfig(color_format='bogus').torgb_out_channels


def test_modconv_tensor_core_shapes():
    """The oracle at the tensor-core shapes of tests/golden/modconv_tc.npz (inputs regenerated from numpy's MT19937 stream)."""
    from oracle.make_golden import MODCONV_TC_CASES, modconv_tc_inputs
    g = load_golden('modconv_tc')
    f4 = O.setup_filter([1, 3, 3, 1])
    for name, (N, cin, cout, H, W, up, demod, has_noise) in MODCONV_TC_CASES.items():
        x, w, s, n = modconv_tc_inputs(name)
        y = O.modulated_conv2d(x, w, s, noise=n, up=up, padding=1, resample_filter=f4, demodulate=demod, flip_weight=(up == 1))
        assert md(y, g[f'{name}_mod']) < 5e-5, name
        if f'{name}_conv' in g:
            y = O.conv2d_resample(x, w, f=(f4 if up > 1 else None), up=up, padding=1, flip_weight=(up == 1))
            assert md(y, g[f'{name}_conv']) < 5e-5, name
            y = O.modulated_conv2d(x, w, s, noise=n, up=up, padding=1, resample_filter=f4, demodulate=demod, flip_weight=(up != 1))
            assert md(y, g[f'{name}_mod_flip']) < 5e-5, name
