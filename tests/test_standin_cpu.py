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


def test_neg_slope_variant_is_refused_by_name(bundles):
    """--neg_slope => conv -> LeakyReLU -> BatchNorm (simple_autoencoder.py:48-53,102-105): folding BN into the conv is not
    exact there, so the hook must say so instead of failing with a KeyError or silently producing wrong features."""
    cfg, ecfg, gp, ep = bundles
    enc = StandInAutoEncoder(ecfg, None, neg_slope=0.2, bn_after_act=True)
    with pytest.raises(RuntimeError, match='batchnorm_after_activation'):
        install.encoder_config_from_reference(enc)
    # default order with a non-default slope is fine and is carried per config
    enc2 = StandInAutoEncoder(ecfg, None, neg_slope=0.2, bn_after_act=False)
    assert install.encoder_config_from_reference(enc2).neg_slope == pytest.approx(0.2)
    # transposed-conv up-sampling (ScaleUpV2, simple_autoencoder.py:130-145)
    enc3 = StandInAutoEncoder(ecfg, None)
    up = enc3.decoder.model[0]
    up.conv = torch.nn.Sequential(torch.nn.ConvTranspose2d(16, 256, 3, stride=2, padding=1, output_padding=1),
                                  torch.nn.LeakyReLU(0.2), torch.nn.BatchNorm2d(256))
    with pytest.raises(RuntimeError, match='ScaleUpV2'):
        install.encoder_config_from_reference(enc3)
