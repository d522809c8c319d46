"""Stand-ins for the reference's engine objects (no /root/reference needed): plain torch modules laid out exactly as the
attributes ``brushstroke_engine_b200.install`` reads -- ``TriadGanPaintEngine`` {G, encoder, device, render_mode, uvs_mapper}
(forger/ui/brush.py:607-805), ``Generator`` (networks_modified.py:228-400) and the sauto ``AutoEncoder``
(simple_autoencoder.py:95-126,155-199,251-297) -- whose ``state_dict()`` keys are the reference's own."""
import types

import torch
import torch.nn as nn


class Single(nn.Module):                                     # SingleConvolution, simple_autoencoder.py:95-109
    def __init__(self, cin, cout, k=3, pad=1, stride=1, neg_slope=None, bn_after_act=False):
        super().__init__()
        conv = nn.Conv2d(cin, cout, k, padding=pad, stride=stride, padding_mode='reflect')
        act = nn.LeakyReLU(inplace=True) if neg_slope is None else nn.LeakyReLU(neg_slope, inplace=True)
        self.conv = nn.Sequential(conv, act, nn.BatchNorm2d(cout)) if bn_after_act else nn.Sequential(conv, nn.BatchNorm2d(cout), act)


class ScaleUp(nn.Module):                                    # simple_autoencoder.py:112-126
    def __init__(self, cin, cout, **kw):
        super().__init__()
        self.up = nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True)
        self.conv = Single(cin, cout, **kw)


class ScaleUpV2(nn.Module):                                  # simple_autoencoder.py:128-148
    def __init__(self, cin, cout, neg_slope):
        super().__init__()
        self.conv = nn.Sequential(nn.ConvTranspose2d(cin, cout, 3, stride=2, padding=1, output_padding=1), nn.LeakyReLU(neg_slope),
                                  nn.BatchNorm2d(cout))


class StandInAutoEncoder(nn.Module):
    """``neg_slope`` + ``bn_after_act=True`` + ``scale_up_v2=True`` is what the reference's ``--neg_slope`` flag builds
    (model_from_flags, simple_autoencoder.py:44-55): the post stages then keep LeakyReLU's default slope (:180-185)."""
    def __init__(self, ecfg, ep=None, scale_up_v2=False, **kw):
        super().__init__()
        enc, dec = nn.Module(), nn.Module()
        chans = [ecfg.pre_filters] + list(ecfg.down_filters) + list(ecfg.post_filters)
        layers = [Single(ecfg.in_channels, ecfg.pre_filters, 7, 3, 1, **kw)]
        for i, c in enumerate(chans[1:]):
            down = i < len(ecfg.down_filters)
            kw_i = kw if down or not scale_up_v2 else {k: v for k, v in kw.items() if k != 'neg_slope'}
            layers.append(Single(chans[i], c, 3, 1, 2 if down else 1, **kw_i))
        enc.model = nn.ModuleList(layers)
        enc.in_channels, enc.num_down_layers = ecfg.in_channels, len(ecfg.down_filters)
        ups, c = [], chans[-1]
        for f in ecfg.up_filters:
            ups.append(ScaleUpV2(c, f, kw.get('neg_slope')) if scale_up_v2 else ScaleUp(c, f, **kw))
            c = f
        dec.model = nn.ModuleList(ups)
        dec.up_layer_filters = list(ecfg.up_filters)
        self.encoder, self.decoder = enc, dec
        self.res = list(ecfg.encode_resolutions)
        self.preproc_name = ecfg.preproc_type
        if ep is not None:
            missing, unexpected = self.load_state_dict(ep, strict=False)
            assert not unexpected, unexpected
        self.eval()


class StandInGenerator:
    """Only what install.generator_config_from_reference / bundle_from_module touch."""
    def __init__(self, cfg, gp):
        self._gp = gp
        self.z_dim, self.w_dim, self.img_resolution, self.img_channels = cfg.z_dim, cfg.w_dim, cfg.img_resolution, cfg.img_channels
        self.mapping = types.SimpleNamespace(num_layers=cfg.mapping_layers)
        last = types.SimpleNamespace(
            conv1=types.SimpleNamespace(weight=gp[f'synthesis.b{cfg.img_resolution}.conv1.weight'], conv_clamp=cfg.conv_clamp),
            torgb=types.SimpleNamespace(color_format=cfg.color_format))
        self.synthesis = types.SimpleNamespace(geom_feature_channels=list(cfg.geom_feature_channels),
                                               geom_feature_resolutions=list(cfg.geom_feature_resolutions))
        setattr(self.synthesis, f'b{cfg.img_resolution}', last)

    def state_dict(self):
        return dict(self._gp)


def standin_engine(cfg, ecfg, gp, ep, device, render_mode='clear', sfactor=None):
    eng = types.SimpleNamespace()
    eng.G = StandInGenerator(cfg, gp)
    eng.encoder = StandInAutoEncoder(ecfg, ep)
    eng.device = torch.device(device)
    eng.render_mode = render_mode
    eng.uvs_mapper = types.SimpleNamespace(get_sfactor=lambda opts: sfactor)
    return eng
