"""CPU: the drop-in hooks read a reference-shaped engine correctly and refuse, by name, the encoder layouts the B200 path
does not build (no /root/reference needed: tests/standin.py mirrors the reference's module layout)."""
import pytest
import torch

from standin import StandInAutoEncoder, StandInGenerator
from brushstroke_engine_b200 import install, params as P


def test_configs_from_stand_in_modules(bundles):
    cfg, ecfg, gp, ep = bundles
    assert install.generator_config_from_reference(StandInGenerator(cfg, gp)) == cfg
    e2 = install.encoder_config_from_reference(StandInAutoEncoder(ecfg, ep))
    assert e2 == ecfg
    b = P.bundle_from_module(StandInAutoEncoder(ecfg, ep))
    for k, v in ep.items():
        if v.dtype.is_floating_point:
            assert k in b and float((b[k] - v).abs().max()) == 0, k


def test_neg_slope_variant_is_recognised(bundles):
    """--neg_slope => conv -> LeakyReLU -> BatchNorm stages + ScaleUpV2 decoder (simple_autoencoder.py:48-53,128-148): the hook
    reads the layout (slope of the pre / down / decoder stages, default slope of the post stages) into the EncoderConfig, and the
    oracle's restatement of that layout equals the torch modules it was read from."""
    from oracle import neube_oracle as O
    cfg, ecfg, gp, ep = bundles
    ecfg2 = P.EncoderConfig(bn_after_activation=True, neg_slope=0.2)
    ep2 = P.init_encoder_params(ecfg2, seed=5, perturb_bn=0.1)
    enc = StandInAutoEncoder(ecfg2, ep2, scale_up_v2=True, neg_slope=0.2, bn_after_act=True)
    got = install.encoder_config_from_reference(enc)
    assert got == ecfg2 and got.bn_after_activation and got.post_neg_slope == pytest.approx(0.01)
    x = torch.rand(2, 1, 64, 64, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        h = x
        for m in enc.encoder.model:
            h = m.conv(h)
        want = [h, enc.decoder.model[0].conv(h)]
        have = O.geometry_encode(ep2, ecfg2, x)
    for w, g in zip(want, have):
        assert float((w - g).abs().max()) < 1e-5


def test_unsupported_encoder_layouts_are_refused_by_name(bundles):
    """Layouts neither model_from_flags branch builds must be named, not fail with a KeyError or silently produce wrong features."""
    cfg, ecfg, gp, ep = bundles
    # BatchNorm after the activation but bilinear ScaleUp decoder stages
    enc = StandInAutoEncoder(ecfg, None, neg_slope=0.2, bn_after_act=True)
    with pytest.raises(RuntimeError, match='batchnorm_after_activation'):
        install.encoder_config_from_reference(enc)
    # default order with a non-default slope is fine and is carried per config
    enc2 = StandInAutoEncoder(ecfg, None, neg_slope=0.2, bn_after_act=False)
    assert install.encoder_config_from_reference(enc2).neg_slope == pytest.approx(0.2)
    # default encoder stages with transposed-conv up-sampling (ScaleUpV2, simple_autoencoder.py:130-145)
    enc3 = StandInAutoEncoder(ecfg, None)
    up = enc3.decoder.model[0]
    up.conv = torch.nn.Sequential(torch.nn.ConvTranspose2d(16, 256, 3, stride=2, padding=1, output_padding=1),
                                  torch.nn.LeakyReLU(0.2), torch.nn.BatchNorm2d(256))
    with pytest.raises(RuntimeError, match='mixed stage layouts'):
        install.encoder_config_from_reference(enc3)
    # a decoder pre layer (--decoder_pre_filters > 0)
    enc4 = StandInAutoEncoder(ecfg, None)
    enc4.decoder.first = torch.nn.Sequential(torch.nn.Conv2d(16, 16, 3))
    with pytest.raises(RuntimeError, match='decoder pre layer'):
        install.encoder_config_from_reference(enc4)
