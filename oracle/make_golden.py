"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference, CPU) on seeded inputs, and check the oracle restatement against
it on the way.  Runs only inside the build container (the GPU box has no
/root/reference); the fixtures it writes are committed.

    python oracle/make_golden.py [--out tests/golden]

TEST INFRASTRUCTURE -- never imported by the product.
"""
from __future__ import annotations

import argparse
import os
import pickle
import sys
import tempfile
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from brushstroke_engine_b200 import params as P            # noqa: E402
from brushstroke_engine_b200 import synthetic               # noqa: E402
from oracle import neube_oracle as O                        # noqa: E402

REF = '/root/reference'


def bootstrap_reference():
    """import the reference with stand-ins for two absent,
    numerically unused imports (skimage.io, matplotlib.pyplot)."""
    from PIL import Image
    sys.path.insert(0, REF)
    stubs = {'skimage': {}, 'skimage.io': {'imread': lambda f, **k: np.array(Image.open(f)),
                                           'imsave': lambda f, a, **k: Image.fromarray(a).save(f)},
             'matplotlib': {}, 'matplotlib.pyplot': {}}
    for name, attrs in stubs.items():
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
    sys.modules['skimage'].io = sys.modules['skimage.io']
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    import thirdparty.stylegan2_ada_pytorch  # noqa: F401  (puts the stylegan root on sys.path)


def build_reference_modules(gp, ep, cfg, ecfg):
    import dnnlib
    from training.networks_modified import Generator
    import forger.experimental.autoenc.factory as factory
    enc_args = argparse.Namespace(model_name='sauto', encoder_in_channels=1, decoder_out_channels=1,
                                  encoder_pre_filters=64, encoder_down_filters='128,256,256',
                                  encoder_post_filters='32,16', decoder_up_filters='256,128,64', neg_slope=None,
                                  decoder_pre_filters=-1, widths='256,128,64', preproc_type=None)
    enc, _ = factory.create_autoencoder(enc_args)
    missing, unexpected = enc.load_state_dict(ep, strict=False)
    assert not unexpected, unexpected
    enc.eval().requires_grad_(False)
    enc.set_default_encode_resolutions([0, 1])
    syn = dnnlib.EasyDict(channel_base=cfg.channel_base, channel_max=cfg.channel_max, num_fp16_res=cfg.num_fp16_res,
                          conv_clamp=cfg.conv_clamp, architecture='orig', color_format=cfg.color_format, color_w_channels=0,
                          enable_geom_linear=False,
                          geom_feature_channels=[enc.feature_channels(r) for r in (0, 1)],
                          geom_feature_resolutions=[enc.featuremap_resolution(128, r) for r in (0, 1)])
    G = Generator(z_dim=cfg.z_dim, c_dim=0, w_dim=cfg.w_dim, img_resolution=cfg.img_resolution, img_channels=3,
                  mapping_kwargs=dnnlib.EasyDict(num_layers=cfg.mapping_layers), synthesis_kwargs=syn)
    missing, unexpected = G.load_state_dict(gp, strict=False)
    assert not unexpected, unexpected
    assert all(('noise_grid' in m or 'resample_filter' in m) for m in missing), missing
    G.eval().requires_grad_(False)
    return G, enc, enc_args


def maxdiff(a, b):
    return float((torch.as_tensor(a).double() - torch.as_tensor(b).double()).abs().max())


def save(out, name, **arrays):
    arrays = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrays.items()}
    path = os.path.join(out, name + '.npz')
    np.savez_compressed(path, **arrays)
    print(f'  wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)')


# --------------------------------------------------------------------------------------

def golden_ops(out):
    from torch_utils.ops import bias_act as rba, upfirdn2d as rup, conv2d_resample as rcr
    from training import networks as rnet
    g = torch.Generator().manual_seed(1234)
    # ---- bias_act: all 9 activations x clamp on/off
    x = torch.randn(2, 5, 6, 7, generator=g) * 3
    b = torch.randn(5, generator=g)
    arrays = {'x': x, 'b': b}
    worst = 0.0
    for act in O.ACTIVATIONS:
        for clamp in (None, 0.7):
            y = rba._bias_act_ref(x, b, dim=1, act=act, clamp=clamp)
            arrays[f'y_{act}_{"c" if clamp else "n"}'] = y
            worst = max(worst, maxdiff(y, O.bias_act(x, b, dim=1, act=act, clamp=clamp)))
    y = rba._bias_act_ref(x, b, dim=1, act='lrelu', alpha=0.1, gain=2.5, clamp=4.0)
    arrays['y_lrelu_custom'] = y
    worst = max(worst, maxdiff(y, O.bias_act(x, b, dim=1, act='lrelu', alpha=0.1, gain=2.5, clamp=4.0)))
    x2 = torch.randn(4, 9, generator=g)
    b2 = torch.randn(9, generator=g)
    arrays['x2'], arrays['b2'] = x2, b2
    arrays['y2_tanh'] = rba._bias_act_ref(x2, b2, dim=1, act='tanh')
    print(f'bias_act: oracle vs reference max diff {worst:.3e}')
    assert worst < 1e-6
    save(out, 'bias_act', **arrays)

    # ---- upfirdn2d
    f4 = rup.setup_filter([1, 3, 3, 1])
    assert maxdiff(f4, O.setup_filter([1, 3, 3, 1])) == 0
    fa = torch.rand(3, 5, generator=g)
    f1 = rup.setup_filter([1, 2, 3, 4, 4, 3, 2, 1])           # separable (>= 8 taps)
    x = torch.randn(2, 3, 9, 11, generator=g)
    cases = {
        'gen':    dict(f=f4, up=1, down=1, padding=[1, 1, 1, 1], flip_filter=False, gain=4.0),   # after transposed conv
        'up2':    dict(f=f4, up=2, down=1, padding=[2, 1, 2, 1], flip_filter=False, gain=4.0),   # upsample2d
        'down2':  dict(f=f4, up=1, down=2, padding=[1, 1, 1, 1], flip_filter=False, gain=1.0),   # downsample2d
        'asym':   dict(f=fa, up=[3, 2], down=[2, 1], padding=[2, -1, 0, 3], flip_filter=True, gain=1.5),
        'neg':    dict(f=fa, up=1, down=1, padding=[-1, 2, -2, 3], flip_filter=False, gain=1.0),
        'sep':    dict(f=f1, up=2, down=1, padding=[4, 3, 4, 3], flip_filter=False, gain=4.0),
        'none':   dict(f=None, up=2, down=1, padding=0, flip_filter=False, gain=1.0),
    }
    arrays = {'x': x, 'f4': f4, 'fa': fa, 'f1': f1}
    worst = 0.0
    for name, kw in cases.items():
        y = rup._upfirdn2d_ref(x, **kw)
        arrays[f'y_{name}'] = y
        worst = max(worst, maxdiff(y, O.upfirdn2d(x, **kw)))
    xg = torch.randn(1, 4, 17, 17, generator=g)               # the generator's (2H+1)^2 -> (2H)^2 case, H = 8
    arrays['xg'] = xg
    arrays['yg'] = rup._upfirdn2d_ref(xg, f4, padding=[1, 1, 1, 1], gain=4.0)
    arrays['y_upsample2d'] = rup.upsample2d(x, f4, impl='ref')
    arrays['y_downsample2d'] = rup.downsample2d(x, f4, impl='ref')
    arrays['y_filter2d'] = rup.filter2d(x, f4, impl='ref')
    worst = max(worst, maxdiff(arrays['y_upsample2d'], O.upsample2d(x, f4)),
                maxdiff(arrays['y_downsample2d'], O.downsample2d(x, f4)),
                maxdiff(arrays['y_filter2d'], O.filter2d(x, f4)))
    print(f'upfirdn2d: oracle vs reference max diff {worst:.3e}')
    assert worst < 2e-6
    save(out, 'upfirdn2d', **arrays)

    # ---- conv2d_resample + modulated_conv2d
    arrays = {}
    worst = 0.0
    for name, (cin, cout, H, up) in {'up1': (6, 5, 8, 1), 'up2': (6, 5, 8, 2), 'up2_odd': (3, 4, 5, 2)}.items():
        x = torch.randn(2, cin, H, H, generator=g)
        w = torch.randn(cout, cin, 3, 3, generator=g)
        s = torch.randn(2, cin, generator=g) + 1
        n = torch.randn(2, 1, H * up, H * up, generator=g)
        arrays[f'{name}_x'], arrays[f'{name}_w'], arrays[f'{name}_s'], arrays[f'{name}_n'] = x, w, s, n
        y = rcr.conv2d_resample(x, w, f=(f4 if up > 1 else None), up=up, padding=1, flip_weight=(up == 1))
        arrays[f'{name}_conv'] = y
        worst = max(worst, maxdiff(y, O.conv2d_resample(x, w, f=(f4 if up > 1 else None), up=up, padding=1,
                                                         flip_weight=(up == 1))))
        for demod in (True, False):
            for fused in (True, False):
                y = rnet.modulated_conv2d(x, w, s, noise=n, up=up, padding=1, resample_filter=f4,
                                          demodulate=demod, flip_weight=(up == 1), fused_modconv=fused)
                if fused:
                    arrays[f'{name}_mod_d{int(demod)}'] = y
                worst = max(worst, maxdiff(y, O.modulated_conv2d(x, w, s, noise=n, up=up, padding=1,
                                                                 resample_filter=f4, demodulate=demod,
                                                                 flip_weight=(up == 1))))
    print(f'conv2d_resample/modulated_conv2d: oracle vs reference max diff {worst:.3e}')
    assert worst < 5e-5
    save(out, 'modconv', **arrays)


MODCONV_TC_CASES = {
    # name: (N, Cin, Cout, H, W, up, demodulate, noise?)   -- shapes the tensor-core kernels take (Cout % 128 == 0, 3x3)
    'c144_up1': (2, 144, 128, 8, 8, 1, True, True),
    'c128_up2': (2, 128, 128, 8, 8, 2, True, True),
    'c384_up2': (1, 384, 128, 5, 6, 2, True, True),
    'c64_o256_up1': (2, 64, 256, 6, 9, 1, True, False),
    'c128_up2_nodemod': (1, 128, 128, 4, 4, 2, False, True),
    'c16_o256_up2': (2, 16, 256, 7, 5, 2, True, False),
}


def modconv_tc_inputs(name):
    """Inputs of one MODCONV_TC_CASES entry from numpy's legacy MT19937 stream (stable across versions and machines), so the
    fixture only has to store the reference's outputs."""
    import zlib
    N, cin, cout, H, W, up, demod, has_noise = MODCONV_TC_CASES[name]
    rs = np.random.RandomState(zlib.crc32(name.encode()) % (2 ** 31))
    x = torch.from_numpy(rs.randn(N, cin, H, W).astype(np.float32))
    w = torch.from_numpy((rs.randn(cout, cin, 3, 3) / np.sqrt(cin * 9)).astype(np.float32))
    s = torch.from_numpy((rs.randn(N, cin) * 0.5 + 1).astype(np.float32))
    n = torch.from_numpy((rs.randn(N, 1, H * up, W * up) * 0.1).astype(np.float32)) if has_noise else None
    return x, w, s, n


def golden_modconv_tc(out):
    """modulated_conv2d / conv2d_resample of the UNMODIFIED reference (CPU, float32) at tensor-core shapes: the
    reference values for the bf16 / fp16 operator-surface path (tests/test_ops_gpu.py::test_modconv_tensor_core_golden)."""
    import training.networks as rnet
    import torch_utils.ops.conv2d_resample as rcr
    f4 = O.setup_filter([1, 3, 3, 1])
    arrays, worst = {}, 0.0
    for name, (N, cin, cout, H, W, up, demod, has_noise) in MODCONV_TC_CASES.items():
        x, w, s, n = modconv_tc_inputs(name)
        fw = up == 1                                                     # SynthesisLayer: flip_weight = (up == 1), networks.py:384
        y = rnet.modulated_conv2d(x, w, s, noise=n, up=up, padding=1, resample_filter=f4, demodulate=demod, flip_weight=fw)
        arrays[f'{name}_mod'] = y
        worst = max(worst, maxdiff(y, O.modulated_conv2d(x, w, s, noise=n, up=up, padding=1, resample_filter=f4,
                                                         demodulate=demod, flip_weight=fw)))
        # the signature's default flip_weight=True on an up-sampling layer (transposed conv with flipped taps), and plain conv2d_resample
        if name in ('c128_up2', 'c64_o256_up1'):
            arrays[f'{name}_mod_flip'] = rnet.modulated_conv2d(x, w, s, noise=n, up=up, padding=1, resample_filter=f4,
                                                               demodulate=demod, flip_weight=not fw)
            arrays[f'{name}_conv'] = rcr.conv2d_resample(x, w, f=(f4 if up > 1 else None), up=up, padding=1, flip_weight=fw)
    print(f'modulated_conv2d (tensor-core shapes): oracle vs reference max diff {worst:.3e}')
    assert worst < 5e-5
    save(out, 'modconv_tc', **arrays)


def golden_generator(out, G, enc, gp, ep, cfg, ecfg):
    z = torch.cat([P.style_z_from_seed(594), P.style_z_from_seed(7)])
    geom = torch.from_numpy(np.concatenate([synthetic.synthetic_patch(128, seed=3, radius=8),
                                            synthetic.synthetic_patch(128, seed=4, radius=3)]))
    gf = enc.encode(geom)
    gfo = O.geometry_encode(ep, ecfg, geom)
    d_enc = max(maxdiff(a, b) for a, b in zip(gf, gfo))
    print(f'encoder: oracle vs reference max diff {d_enc:.3e}')
    assert d_enc < 2e-5
    pos = torch.tensor([[88, 176], [1144, 264]])
    arrays = {'z': z, 'geom': geom, 'positions': pos, 'g0': gf[0], 'g1_sub': gf[1][:, ::8]}
    for tag, positions in (('nopos', None), ('pos', pos)):
        img32, dbg32 = G(z, None, gf, positions=positions, return_debug_data=True, return_features=[64],
                         noise_mode='const', force_fp32=True)
        img16, dbg16 = G(z, None, gf, positions=positions, return_debug_data=True, noise_mode='const')
        imgo, dbgo = O.generator_forward(gp, cfg, z, gf, positions=positions, return_features=[64])
        d = max(maxdiff(img32, imgo), maxdiff(dbg32['uvs'], dbgo['uvs']), maxdiff(dbg32['colors'], dbgo['colors']))
        dws = maxdiff(dbg32['ws'], dbgo['ws'])
        df = maxdiff(dbg32['features64'], dbgo['features64'])
        print(f'generator[{tag}]: oracle vs reference fp32 max diff img/uvs/colors {d:.3e}, ws {dws:.3e}, '
              f'features64 {df:.3e}; reference fp16 vs fp32 {maxdiff(img16, img32):.3e}')
        assert d < 1e-4 and dws < 1e-5
        arrays[f'img32_{tag}'] = img32
        arrays[f'uvs32_{tag}'] = dbg32['uvs']
        arrays[f'colors32_{tag}'] = dbg32['colors']
        arrays[f'img16_{tag}'] = img16.to(torch.float16)
        arrays[f'feat64_sub_{tag}'] = dbg32['features64'][:, ::16, ::2, ::2]
        arrays['ws'] = dbg32['ws']
    # shifted-noise closed form vs grid_sample
    nc = gp['synthesis.b16.conv1.noise_const']
    npos = (pos % 128) / 127
    dn = maxdiff(O.shifted_noise(nc, npos), O.shifted_noise_closed_form(nc, npos))
    print(f'shifted noise closed form vs grid_sample: {dn:.3e}')
    assert dn < 1e-5
    arrays['noise16_pos'] = O.shifted_noise(nc, npos)
    arrays['gen_digest'] = np.frombuffer(P.bundle_digest(gp).encode(), dtype=np.uint8)
    arrays['enc_digest'] = np.frombuffer(P.bundle_digest(ep).encode(), dtype=np.uint8)
    save(out, 'generator', **arrays)


def reference_engine(G, enc, enc_args, cfg):
    """The reference's own ``PaintEngineFactory`` on a synthetic snapshot, CPU."""
    import forger.ui.brush as brush
    snap = {'G': G, 'D': torch.nn.Identity(), 'G_ema': G, 'training_set_kwargs': None, 'augment_pipe': None,
            'args': argparse.Namespace(color_format='triad', geom_inject_resolutions=[0, 1]),
            'encoder': {'args': enc_args, 'model_state': enc.state_dict()}}
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, 'snapshot.pkl')
        with open(path, 'wb') as f:
            pickle.dump(snap, f)
        engine = brush.PaintEngineFactory.create(gan_checkpoint=path, device=torch.device('cpu'))
    # make the engine run the fp32 path so the fixture pins exact semantics (the stock engine is mixed fp16)
    for res in cfg.block_resolutions:
        getattr(engine.G.synthesis, f'b{res}').use_fp16 = False
    return engine


def wire_script():
    """The client side of the recorded session: JSON strings and binary render requests, in order (shared with the tests
    only through the fixture it produces)."""
    from brushstroke_engine_b200 import server

    def stroke(seed, radius=6):
        g = synthetic.synthetic_patch(128, seed=seed, radius=radius)[0, 0]              # float, 0 = stroke
        rgba = np.zeros((128, 128, 4), dtype=np.uint8)
        rgba[..., 3] = np.round((1 - g) * 255).astype(np.uint8)
        rgba[..., 0] = 30                                                                # ignored by the server
        return rgba
    import json
    msgs = [json.dumps({'type': 'set_brush', 'seed': 594}),
            json.dumps({'type': 'new_canvas', 'rows': 300, 'cols': 420, 'feature_blending': 2}),
            json.dumps({'type': 'set_option', 'option': 'positions', 'value': True}),
            server.encode_render_request(stroke(21), 0, 0, 10),
            server.encode_render_request(stroke(22), 89, 1, 10, colors=[(1, 255, 0, 128)], extra_data=5),    # snapped to (88, 0)
            server.encode_render_request(stroke(23), 44, 61, 10, colors=[(0, 10, 200, 30), (2, 250, 250, 240)]),
            server.encode_render_request(stroke(24), 176, 88, 0, debug=False),
            json.dumps({'type': 'set_render_mode', 'mode': 'full'}),
            server.encode_render_request(stroke(25), 130, 30, 10),
            json.dumps({'type': 'set_render_mode', 'mode': 'clear'}),
            json.dumps({'type': 'set_option', 'option': 'uvs_mapping', 'value': True}),
            server.encode_render_request(stroke(26), 200, 120, 10),
            json.dumps({'type': 'set_option', 'option': 'uvs_mapping', 'value': False}),
            json.dumps({'type': 'set_option', 'option': 'positions', 'value': False}),
            json.dumps({'type': 'new_canvas', 'rows': 300, 'cols': 420, 'feature_blending': 0}),
            server.encode_render_request(stroke(27), 7, 9, 10, extra_data=3),
            json.dumps({'type': 'set_brush'}),                                            # random seed from the session's rng
            server.encode_render_request(stroke(28), 64, 64, 0),
            b'\x00\x00',                                                               # undecodable: dropped without an answer
            json.dumps({'type': 'bogus'})]
    return msgs


def golden_wire(out, G, enc, enc_args, cfg):
    """A whole interactive session through the reference's own ``DrawingWebSocketHandler`` (forger/ui/util.py:107-245;
    Tornado replaced by a stand-in base class -- the handler's logic is untouched): every message the client sends and
    every message the handler writes back."""
    wh = types.ModuleType('tornado.websocket')
    wh.WebSocketHandler = type('WebSocketHandler', (), {})
    gen = types.ModuleType('tornado.gen')
    gen.coroutine = lambda f: f
    tor = types.ModuleType('tornado')
    tor.websocket, tor.gen = wh, gen
    sys.modules.update({'tornado': tor, 'tornado.websocket': wh, 'tornado.gen': gen})
    import forger.ui.util as ui_util
    engine = reference_engine(G, enc, enc_args, cfg)
    geo5 = torch.from_numpy(np.concatenate([synthetic.synthetic_patch(128, seed=10 + i, radius=8) for i in range(5)]))
    geo5_thick = torch.from_numpy(np.concatenate([synthetic.synthetic_patch(128, seed=10 + i, radius=12) for i in range(5)]))
    engine.uvs_mapper.geom_feature = engine.encoder.encode(geo5)
    engine.uvs_mapper.fmask = geo5 < 0.01
    engine.uvs_mapper.bmask = geo5_thick > 0.99
    handler = ui_util.DrawingWebSocketHandler()
    handler.initialize(engine, style_seed=3, debug_dir=None, saved_zs_filename=None, libraries={})
    written = []
    handler.write_message = lambda m, binary=False: written.append((m, binary))
    arrays = {}

    def record(tag):
        for j, (m, binary) in enumerate(written):
            import json
            arrays[f'{tag}_out{j}'] = np.frombuffer(m if binary else json.dumps(m).encode(), dtype=np.uint8)
            arrays[f'{tag}_out{j}_binary'] = np.uint8(binary)
        arrays[f'{tag}_nout'] = np.int32(len(written))
        written.clear()
    with torch.no_grad():
        handler.open()
        record('open')
        msgs = wire_script()
        for i, m in enumerate(msgs):
            handler.on_message(m)
            arrays[f'm{i}_in'] = np.frombuffer(m if isinstance(m, bytes) else m.encode(), dtype=np.uint8)
            arrays[f'm{i}_in_binary'] = np.uint8(isinstance(m, bytes))
            record(f'm{i}')
    arrays['n_msgs'] = np.int32(len(msgs))
    n_bin = sum(int(arrays[f'm{i}_nout']) for i in range(len(msgs)) if arrays[f'm{i}_in_binary'])
    print(f'wire: {len(msgs)} client messages, {n_bin} binary answers recorded')
    assert n_bin == 8
    save(out, 'wire', **arrays)


def golden_engine(out, G, enc, enc_args, gp, ep, cfg, ecfg):
    """Engine composite + stylizer loop through the reference's own PaintingHelper."""
    import forger.ui.brush as brush
    import forger.viz.style_transfer as style_transfer
    engine = reference_engine(G, enc, enc_args, cfg)

    guidance = synthetic.synthetic_guidance(300, 260, num_lines=10, seed=5, radii=(1, 3, 9))
    crop_margin = 10
    arrays = {'guidance': guidance}
    sfactor = None
    for level in (0, 2):
        for mode in ('clear', 'full'):
            if level == 2 and mode == 'full':
                continue
            helper = brush.PaintingHelper(engine)
            opts = brush.GanBrushOptions()
            opts.set_style(P.style_z_from_seed(594), '594')
            geom = O.pad_geo(guidance, crop_margin)
            crops, geom = style_transfer.generate_stitching_crops(geom, 128, mode='all', overlap_margin=crop_margin * 2)
            crops_o, geom_o = O.generate_stitching_crops(O.pad_geo(guidance, crop_margin), 128, 'all', crop_margin * 2)
            assert crops == crops_o and np.array_equal(geom, geom_o)
            result = np.zeros((geom.shape[0], geom.shape[1], 4), dtype=np.uint8)
            helper.make_new_canvas(result.shape[0], result.shape[1], feature_blending=level)
            helper.set_render_mode(mode)
            metas = []
            with torch.no_grad():
                for (y, x, _, _) in crops:
                    opts.set_position(x, y)
                    gpatch = 255 - geom[y:y + 128, x:x + 128, :]
                    res, _, meta = helper.render_stroke(gpatch, None, opts, meta={'x': x, 'y': y, 'crop_margin': crop_margin})
                    result[meta['y']:meta['y'] + res.shape[0], meta['x']:meta['x'] + res.shape[1], :] = res
                    metas.append((meta['y'], meta['x']))
            final = result[crop_margin:crop_margin + guidance.shape[0], crop_margin:crop_margin + guidance.shape[1], :]
            canvas_o, _, metas_o = O.stylize(gp, ep, cfg, ecfg, guidance, P.style_z_from_seed(594), crop_margin,
                                            'all', mode, level)
            assert metas == metas_o
            diff = np.abs(final.astype(np.int32) - canvas_o.astype(np.int32))
            print(f'stylize[level={level},{mode}]: {len(crops)} crops, oracle vs reference uint8 max diff {diff.max()}, '
                  f'differing bytes {int((diff > 0).sum())}/{diff.size}')
            assert diff.max() <= 1
            arrays[f'canvas_l{level}_{mode}'] = final
            arrays['crops'] = np.array(crops, dtype=np.int32)
            arrays['metas'] = np.array(metas, dtype=np.int32)
    # tile ownership closed form == loop
    H, W = geom.shape[:2]
    owner = np.full((H, W), -1, dtype=np.int32)
    for i, (y, x, _, _) in enumerate(crops):
        owner[y + crop_margin:y + 128 - crop_margin, x + crop_margin:x + 128 - crop_margin] = i
    nrows = (guidance.shape[0] + crop_margin) // 88 + 1
    ncols = (guidance.shape[1] + crop_margin) // 88 + 1
    Y, X = np.mgrid[0:H, 0:W]
    r, c = O.tile_owner_closed_form(Y, X, crop_margin, 88, nrows, ncols)
    closed = np.where((Y >= crop_margin) & (X >= crop_margin) & (Y < (nrows - 1) * 88 + 128 - crop_margin)
                      & (X < (ncols - 1) * 88 + 128 - crop_margin), r * ncols + c, -1)
    assert np.array_equal(owner, closed), 'tile ownership closed form mismatch'
    print('tile ownership closed form == placement loop')

    # UVS mapping on reference code with our own synthetic geometry (mapper.py:53-72,117-135)
    from forger.ui.mapper import StyleUVSMapper
    geo5 = torch.from_numpy(np.concatenate([synthetic.synthetic_patch(128, seed=10 + i, radius=8) for i in range(5)]))
    geo5_thick = torch.from_numpy(np.concatenate([synthetic.synthetic_patch(128, seed=10 + i, radius=12) for i in range(5)]))
    mapper = engine.uvs_mapper
    mapper.geom_feature = engine.encoder.encode(geo5)
    mapper.fmask = geo5 < 0.01
    mapper.bmask = geo5_thick > 0.99
    opts = brush.GanBrushOptions()
    opts.set_style(P.style_z_from_seed(594), '594')
    sf = mapper.get_sfactor(opts)
    _, raw = mapper._render(opts, mapper.geom_feature)
    sfo = O.uvs_sfactor(raw['uvs'][:, 2:3].float(), mapper.bmask)
    mapped = StyleUVSMapper._map_style_s(sf, raw['uvs'].float().clone())
    mapped_o = O.map_style_s(sfo, raw['uvs'].float())
    print(f'uvs mapper: sfactor ref {float(sf):.6f} oracle {float(sfo):.6f}; mapped max diff {maxdiff(mapped, mapped_o):.3e}')
    assert abs(float(sf) - float(sfo)) < 1e-5 and maxdiff(mapped, mapped_o) < 1e-5
    arrays['uvs5_sub'] = raw['uvs'].float()[:, :, ::2, ::2]
    arrays['mapped5_sub'] = mapped[:, :, ::2, ::2]
    arrays['sfactor'] = np.float32(float(sf))
    save(out, 'engine', **arrays)


def golden_canvas(out, ep, ecfg):
    """The 'canvas' colour format (ToRGBColorTriadLayer with 3 + 5 outputs, networks.py:433-481) through the reference
    Generator and the reference CanvasPaintEngine (brush.py:870-935) in all four render modes."""
    import forger.ui.brush as brush
    cfg = P.GeneratorConfig(color_format='canvas')
    gp = P.init_generator_params(cfg, seed=3, perturb=0.1)
    G, enc, enc_args = build_reference_modules(gp, ep, cfg, ecfg)
    z = torch.cat([P.style_z_from_seed(594), P.style_z_from_seed(7)])
    geom = torch.from_numpy(np.concatenate([synthetic.synthetic_patch(128, seed=3, radius=8),
                                            synthetic.synthetic_patch(128, seed=4, radius=3)]))
    gf = enc.encode(geom)
    img32, dbg = G(z, None, gf, return_debug_data=True, noise_mode='const', force_fp32=True)
    imgo, dbgo = O.generator_forward(gp, cfg, z, gf)
    d = max(maxdiff(img32, imgo), *[maxdiff(dbg[k], dbgo[k]) for k in ('uvs', 'colors', 'canvas', 'alpha_fg', 'alpha')])
    print(f'canvas format generator: oracle vs reference fp32 max diff {d:.3e}')
    assert d < 1e-4
    arrays = {'z': z, 'geom': geom, 'img32': img32, 'uvs32_sub': dbg['uvs'][:, :, ::2, ::2], 'colors32': dbg['colors'],
              'canvas32_sub': dbg['canvas'][:, :, ::2, ::2], 'alpha32_sub': dbg['alpha'][:, :, ::2, ::2], 'gen_digest': np.frombuffer(P.bundle_digest(gp).encode(), dtype=np.uint8)}
    snap = {'G': G, 'D': torch.nn.Identity(), 'G_ema': G, 'training_set_kwargs': None, 'augment_pipe': None,
            'args': argparse.Namespace(color_format='canvas', geom_inject_resolutions=[0, 1]),
            'encoder': {'args': enc_args, 'model_state': enc.state_dict()}}
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, 'snapshot.pkl')
        with open(path, 'wb') as f:
            pickle.dump(snap, f)
        engine = brush.PaintEngineFactory.create(gan_checkpoint=path, device=torch.device('cpu'))
    assert type(engine).__name__ == 'CanvasPaintEngine'
    for res in cfg.block_resolutions:
        getattr(engine.G.synthesis, f'b{res}').use_fp16 = False
    opts = brush.GanBrushOptions()
    opts.set_style(z)
    opts.set_color(1, np.array([255, 0, 128], dtype=np.uint8))
    for mode in ('clear', 'stroke', 'canvas', 'full'):
        engine.set_render_mode(mode)
        rgba, raw, _ = engine._render_stroke_torch(geom, None, opts)
        ref = O.canvas_composite(dbgo['uvs'], dbgo['colors'], dbgo['alpha_fg'], dbgo['canvas'], mode,
                                 color1=torch.tensor([1.0, 0.0, 128 / 255]))
        dm = maxdiff(rgba, ref)
        print(f'canvas engine [{mode}]: oracle vs reference max diff {dm:.3e}')
        assert dm < 1e-4
        arrays[f'rgba_{mode}_sub'] = rgba[:, :, ::2, ::2]
    save(out, 'canvas', **arrays)


def golden_encoder_v2(out):
    """The ``--neg_slope`` autoencoder (simple_autoencoder.py:48-53: conv -> LeakyReLU -> BatchNorm stages, ``ScaleUpV2``
    transposed-conv decoder) built by the reference's own factory with neg_slope = 0.2; oracle restatement checked on the way."""
    import forger.experimental.autoenc.factory as factory
    ecfg = P.EncoderConfig(bn_after_activation=True, neg_slope=0.2)
    ep = P.init_encoder_params(ecfg, seed=5, perturb_bn=0.1)
    enc_args = argparse.Namespace(model_name='sauto', encoder_in_channels=1, decoder_out_channels=1,
                                  encoder_pre_filters=64, encoder_down_filters='128,256,256',
                                  encoder_post_filters='32,16', decoder_up_filters='256,128,64', neg_slope=0.2,
                                  decoder_pre_filters=-1, widths='256,128,64', preproc_type=None)
    enc, _ = factory.create_autoencoder(enc_args)
    missing, unexpected = enc.load_state_dict(ep, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith('decoder.model.') and int(k.split('.')[2]) >= 1 or 'num_batches_tracked' in k for k in missing), missing
    enc.eval().requires_grad_(False)
    enc.set_default_encode_resolutions([0, 1])
    geom = torch.from_numpy(np.concatenate([synthetic.synthetic_patch(128, seed=3, radius=8),
                                            synthetic.synthetic_patch(128, seed=4, radius=3)]))
    gf = enc.encode(geom)
    gfo = O.geometry_encode(ep, ecfg, geom)
    d_enc = max(maxdiff(a, b) for a, b in zip(gf, gfo))
    print(f'encoder (--neg_slope variant): oracle vs reference max diff {d_enc:.3e}')
    assert d_enc < 2e-5
    save(out, 'encoder_v2', geom=geom, g0=gf[0], g1_sub=gf[1][:, ::8],
         enc_digest=np.frombuffer(P.bundle_digest(ep).encode(), dtype=np.uint8))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=os.path.join(REPO, 'tests', 'golden'))
    ap.add_argument('--only', default=None, help="write one fixture only: 'modconv_tc' | 'wire' | 'encoder_v2'")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    torch.set_grad_enabled(False)
    bootstrap_reference()
    if args.only == 'modconv_tc':
        return golden_modconv_tc(args.out)
    if args.only == 'encoder_v2':
        return golden_encoder_v2(args.out)
    cfg, ecfg = P.GeneratorConfig(), P.EncoderConfig()
    gp = P.init_generator_params(cfg, seed=0, perturb=0.1)
    ep = P.init_encoder_params(ecfg, seed=1, perturb_bn=0.1)
    G, enc, enc_args = build_reference_modules(gp, ep, cfg, ecfg)
    if args.only == 'wire':
        return golden_wire(args.out, G, enc, enc_args, cfg)
    golden_ops(args.out)
    golden_generator(args.out, G, enc, gp, ep, cfg, ecfg)
    golden_engine(args.out, G, enc, enc_args, gp, ep, cfg, ecfg)
    golden_canvas(args.out, ep, ecfg)
    golden_modconv_tc(args.out)
    golden_wire(args.out, G, enc, enc_args, cfg)
    golden_encoder_v2(args.out)


if __name__ == '__main__':
    main()
