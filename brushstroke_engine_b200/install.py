"""Drop the B200 path into a running reference (NeuBE) process.

The reference resolves its ops through module attributes, so a maintainer can switch the hot path without touching
the reference tree (SURVEY.md section 8b "Installation point"):

* ``install_ops()`` rebinds the operator surface on the *singleton* modules ``torch_utils.ops.{bias_act,upfirdn2d,
  conv2d_resample}`` (pickle-embedded layer code resolves them from ``sys.modules``,
  SG2/torch_utils/persistence.py:216-227) and ``modulated_conv2d`` on both import aliases of ``training.networks``
  (SG2/__init__.py:1-5 makes two module objects).  Forward-only: autograd callers must call ``uninstall_ops()`` first.
* ``engine_from_reference(engine)`` builds a ``TriadPaintEngine`` from a reference ``TriadGanPaintEngine`` (weights are
  read through ``state_dict()``), and ``attach_fast_path(engine)`` replaces its ``_render_stroke_torch`` so that
  ``forger.viz.paint_image_main`` / ``forger.ui`` keep calling ``PaintingHelper.render_stroke`` unchanged.
"""
from __future__ import annotations

import sys
import types

from . import bias_act as _ba
from . import conv2d_resample as _cr
from . import modconv as _mc
from . import upfirdn2d as _up
from .params import EncoderConfig, GeneratorConfig, bundle_from_module

_saved = {}

_OPS = {
    'torch_utils.ops.bias_act': {'bias_act': _ba.bias_act},
    'torch_utils.ops.upfirdn2d': {'upfirdn2d': _up.upfirdn2d, 'filter2d': _up.filter2d, 'upsample2d': _up.upsample2d,
                                  'downsample2d': _up.downsample2d},
    'torch_utils.ops.conv2d_resample': {'conv2d_resample': _cr.conv2d_resample},
    'torch_utils.ops.fma': {'fma': _mc.fma},
    'training.networks': {'modulated_conv2d': _mc.modulated_conv2d},
    'thirdparty.stylegan2_ada_pytorch.training.networks': {'modulated_conv2d': _mc.modulated_conv2d},
}


def install_ops():
    """Rebind the reference's operator surface to the CUDA implementations.  Returns the list of patched names."""
    patched = []
    for mod_name, attrs in _OPS.items():
        mod = sys.modules.get(mod_name)
        if mod is None:
            continue
        for name, fn in attrs.items():
            if hasattr(mod, name):
                _saved.setdefault((mod_name, name), getattr(mod, name))
                setattr(mod, name, fn)
                patched.append(f'{mod_name}.{name}')
    if not patched:
        raise RuntimeError('install_ops: the reference is not imported (import thirdparty.stylegan2_ada_pytorch first)')
    return patched


def uninstall_ops():
    for (mod_name, name), fn in list(_saved.items()):
        mod = sys.modules.get(mod_name)
        if mod is not None:
            setattr(mod, name, fn)
        del _saved[(mod_name, name)]


def generator_config_from_reference(G) -> GeneratorConfig:
    syn = G.synthesis
    last = getattr(syn, f'b{G.img_resolution}')
    c128 = last.conv1.weight.shape[0]
    return GeneratorConfig(z_dim=G.z_dim, w_dim=G.w_dim, img_resolution=G.img_resolution, img_channels=G.img_channels,
                           mapping_layers=G.mapping.num_layers, channel_base=c128 * G.img_resolution, channel_max=c128,
                           conv_clamp=last.conv1.conv_clamp,
                           geom_feature_channels=tuple(syn.geom_feature_channels),
                           geom_feature_resolutions=tuple(syn.geom_feature_resolutions),
                           color_format=getattr(last.torgb, 'color_format', 'triad'))


def _encoder_layer_order(enc) -> dict:
    """Read the stage layout of a reference ``sauto`` autoencoder.  Two layouts exist (simple_autoencoder.py:48-53):
      * default: every stage conv -> BatchNorm -> LeakyReLU with one slope, bilinear ``ScaleUp`` decoder stages (eval-mode
        BatchNorm folds into the preceding convolution);
      * ``--neg_slope``: conv -> LeakyReLU -> BatchNorm (``batchnorm_after_activation=True``), the pre / down stages and the
        ``ScaleUpV2`` (ConvTranspose2d) decoder stages with the flag's slope, the post stages with LeakyReLU's default
        (simple_autoencoder.py:180-185) -- BatchNorm folds FORWARD into the next convolution (``GeometryEncoder._build_v2``).
    Anything else (mixed orders, a decoder ``first`` layer, unknown stage modules) is refused by name instead of producing
    wrong features.  Returns the EncoderConfig fields that describe the layout."""
    import torch.nn as nn
    if getattr(enc.decoder, 'first', None) is not None:
        raise RuntimeError('encoder_config_from_reference: unsupported geometry encoder: decoder pre layer (--decoder_pre_filters > 0, '
                           'simple_autoencoder.py:217-231)')
    n_down = enc.encoder.num_down_layers
    n_feat = max(enc.res) if isinstance(enc.res, (list, tuple)) else int(enc.res)
    dec_stages = [m for m in enc.decoder.model if hasattr(m, 'conv')][:max(n_feat, 0)]
    orders, slopes = [], []
    for m in list(enc.encoder.model) + dec_stages:
        transposed = isinstance(m.conv, nn.Sequential) and isinstance(m.conv[0], nn.ConvTranspose2d)        # ScaleUpV2
        single = m.conv if isinstance(m.conv, nn.Sequential) else m.conv.conv       # SingleConvolution / ScaleUp(SingleConvolution)
        kinds = [type(x) for x in single]
        if len(kinds) != 3 or not issubclass(kinds[0], (nn.Conv2d, nn.ConvTranspose2d)):
            raise RuntimeError(f'encoder_config_from_reference: unsupported encoder stage {[k.__name__ for k in kinds]}')
        if issubclass(kinds[1], nn.BatchNorm2d) and issubclass(kinds[2], nn.LeakyReLU) and not transposed:
            orders.append('bn_act'); slopes.append(float(single[2].negative_slope))
        elif issubclass(kinds[1], nn.LeakyReLU) and issubclass(kinds[2], nn.BatchNorm2d):
            orders.append('act_bn_T' if transposed else 'act_bn'); slopes.append(float(single[1].negative_slope))
            if transposed:
                c = single[0]
                if (tuple(c.kernel_size), tuple(c.stride), tuple(c.padding), tuple(c.output_padding), tuple(c.dilation)) != \
                        ((3, 3), (2, 2), (1, 1), (1, 1), (1, 1)):
                    raise RuntimeError('encoder_config_from_reference: unsupported ScaleUpV2 geometry (expected ConvTranspose2d 3x3, stride 2, '
                                       'padding 1, output_padding 1, simple_autoencoder.py:133-141)')
        else:
            raise RuntimeError(f'encoder_config_from_reference: unsupported encoder stage {[k.__name__ for k in kinds]}')
    n_enc = len(enc.encoder.model)
    if all(o == 'bn_act' for o in orders):
        if len(set(slopes)) != 1:
            raise RuntimeError(f'encoder_config_from_reference: layers use different LeakyReLU slopes {sorted(set(slopes))}')
        return dict(neg_slope=slopes[0])
    if all(o == 'act_bn' for o in orders[:n_enc]) and all(o == 'act_bn_T' for o in orders[n_enc:]):
        main, post = set(slopes[:1 + n_down] + slopes[n_enc:]), set(slopes[1 + n_down:n_enc])
        if len(main) != 1 or len(post) > 1:
            raise RuntimeError(f'encoder_config_from_reference: layers use different LeakyReLU slopes {sorted(main)} / {sorted(post)}')
        return dict(neg_slope=main.pop(), bn_after_activation=True, post_neg_slope=post.pop() if post else 0.01)
    raise RuntimeError(f'encoder_config_from_reference: unsupported geometry encoder: mixed stage layouts {orders} (expected the default '
                       'sauto layout or the --neg_slope variant with batchnorm_after_activation=True and ScaleUpV2, '
                       'simple_autoencoder.py:48-53)')


def encoder_config_from_reference(enc) -> EncoderConfig:
    e, d = enc.encoder, enc.decoder
    layout = _encoder_layer_order(enc)
    convs = [m.conv[0] for m in e.model]
    n_down = e.num_down_layers
    return EncoderConfig(in_channels=e.in_channels, pre_filters=convs[0].out_channels,
                         down_filters=tuple(c.out_channels for c in convs[1:1 + n_down]),
                         post_filters=tuple(c.out_channels for c in convs[1 + n_down:]),
                         up_filters=tuple(d.up_layer_filters), encode_resolutions=tuple(enc.res),
                         preproc_type=enc.preproc_name, **layout)


def engine_from_reference(ref_engine, mode: str = 'bf16'):
    """reference ``TriadGanPaintEngine`` / ``CanvasPaintEngine`` -> ``TriadPaintEngine`` / ``CanvasPaintEngine`` sharing its
    weights, device and render mode (the class follows the generator's colour format, as brush.py:594-599)."""
    from .engine import TriadPaintEngine, CanvasPaintEngine
    gp = bundle_from_module(ref_engine.G)
    ep = bundle_from_module(ref_engine.encoder)
    gen_cfg = generator_config_from_reference(ref_engine.G)
    cls = CanvasPaintEngine if gen_cfg.color_format == 'canvas' else TriadPaintEngine
    eng = cls(gp, ep, ref_engine.device, mode=mode, gen_cfg=gen_cfg, enc_cfg=encoder_config_from_reference(ref_engine.encoder))
    eng.set_render_mode(ref_engine.render_mode)
    return eng


def attach_fast_path(ref_engine, mode: str = 'bf16'):
    """Route a reference engine's ``_render_stroke_torch`` through the B200 engine (keeps ``render_mode`` in sync and
    re-uses the reference's own UVS-mapper factors, which depend on its bundled geometry images)."""
    fast = engine_from_reference(ref_engine, mode)

    def _render_stroke_torch(self, geom, canvas, opts, **generator_kwargs):
        fast.render_mode = self.render_mode
        if opts.enable_uvs_mapping:
            sid = opts.style_id
            fast.uvs_mapper.sfactors[sid] = self.uvs_mapper.get_sfactor(opts)
        return fast._render_stroke_torch(geom, canvas, opts, **generator_kwargs)

    ref_engine._nbe_fast = fast
    ref_engine._render_stroke_torch = types.MethodType(_render_stroke_torch, ref_engine)
    return fast
