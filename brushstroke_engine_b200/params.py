"""Parameter bundles for the NeuBE generator-forward hot path.

A *bundle* is a flat ``dict[str, torch.Tensor]`` (CPU, float32) keyed by the
reference's own ``state_dict`` names, so that one bundle can be

* loaded into the reference modules (``G.load_state_dict(bundle, strict=False)``;
  used only by ``oracle/make_golden.py`` inside the build container),
* consumed by the oracle restatement (``oracle/neube_oracle.py``), and
* compiled into a device-side plan by :mod:`brushstroke_engine_b200.generator`.

Reference naming / shapes (checked against the reference in the build container):
``synthesis.b{res}.conv{0,1}.{weight,bias,noise_const,noise_strength,affine.weight,affine.bias}``,
``synthesis.b4.const``, ``synthesis.b{R}.torgb.{weight,bias,color_bias,affine.weight,affine.bias}``,
``mapping.fc{i}.{weight,bias}``, ``mapping.w_avg``
(thirdparty/stylegan2_ada_pytorch/training/networks.py:93-122,215-290,303-391,416-485,540-628;
networks_modified.py:42-118).  Encoder names follow
forger/experimental/autoenc/simple_autoencoder.py:95-199,251-297.

The random initialisation below draws from the same *distributions* as the
reference constructors (randn weights, zero biases, bias_init=1 affines,
``randn/lr_multiplier`` mapping weights, xavier-normal encoder convs) but in our
own order from our own ``torch.Generator`` -- the GPU box has no reference, so
"random-init weights of the named architecture" must be reproducible from a
seed alone.
"""
from __future__ import annotations

import dataclasses
import hashlib
import math
from typing import Dict, List, Sequence

import numpy as np
import torch

Bundle = Dict[str, torch.Tensor]


@dataclasses.dataclass(frozen=True)
class GeneratorConfig:
    """The "style2 architecture" of SURVEY.md section 5 (train_flags.txt:1-20)."""
    z_dim: int = 64
    w_dim: int = 64
    img_resolution: int = 128
    img_channels: int = 3
    mapping_layers: int = 4
    mapping_lr_multiplier: float = 0.01
    channel_base: int = 16384
    channel_max: int = 128
    num_fp16_res: int = 4
    conv_clamp: float = 256.0
    geom_feature_channels: Sequence[int] = (16, 256)
    geom_feature_resolutions: Sequence[int] = (16, 32)
    color_format: str = 'triad'                 # 'triad' (stock models) or 'canvas' (ToRGBColorTriadLayer, networks.py:433-440)

    @property
    def block_resolutions(self) -> List[int]:
        log2 = int(math.log2(self.img_resolution))
        return [2 ** i for i in range(2, log2 + 1)]

    def channels(self, res: int) -> int:
        return min(self.channel_base // res, self.channel_max)

    def block_in_channels(self, res: int) -> int:
        """Input channels of ``b{res}.conv0`` incl. injected geometry
        (networks_modified.py:83-92)."""
        if res == 4:
            return 0
        c = self.channels(res // 2)
        if (res // 2) in self.geom_feature_resolutions:
            c += self.geom_feature_channels[list(self.geom_feature_resolutions).index(res // 2)]
        return c

    @property
    def torgb_out_channels(self) -> int:
        """3 UVS maps, plus a 3-channel canvas and 2 alpha logits for the 'canvas' format (networks.py:434-440)."""
        if self.color_format not in ('triad', 'canvas'):
            raise RuntimeError(f'Unknown format {self.color_format}')
        return self.img_channels + (5 if self.color_format == 'canvas' else 0)

    @property
    def num_ws(self) -> int:
        # one per conv layer + the final torgb (networks_modified.py:113-116)
        return 2 * len(self.block_resolutions) - 1 + 1

    def layer_names(self) -> List[str]:
        names = []
        for res in self.block_resolutions:
            if res > 4:
                names.append(f'b{res}.conv0')
            names.append(f'b{res}.conv1')
        return names

    def fp16_resolution(self) -> int:
        log2 = int(math.log2(self.img_resolution))
        return max(2 ** (log2 + 1 - self.num_fp16_res), 8)


@dataclasses.dataclass(frozen=True)
class EncoderConfig:
    """``sauto`` geometry autoencoder (simple_autoencoder.py:155-199, model_from_flags); only the encode path is described.
    Default flags: every stage is conv -> BatchNorm -> LeakyReLU(0.01) and the decoder up-samples bilinearly (``ScaleUp``).
    ``bn_after_activation=True`` is the ``--neg_slope`` variant (simple_autoencoder.py:48-53): conv -> LeakyReLU -> BatchNorm,
    the pre / down layers and the decoder use ``neg_slope``, the post layers keep ``post_neg_slope`` (their constructor gets no
    slope, :180-185), and the decoder stages are ``ScaleUpV2`` (ConvTranspose2d 3x3 stride 2, :128-148)."""
    in_channels: int = 1
    pre_filters: int = 64
    down_filters: Sequence[int] = (128, 256, 256)
    post_filters: Sequence[int] = (32, 16)
    up_filters: Sequence[int] = (256, 128, 64)
    encode_resolutions: Sequence[int] = (0, 1)
    preproc_type: str | None = None
    bn_eps: float = 1e-5
    neg_slope: float = 0.01
    bn_after_activation: bool = False
    post_neg_slope: float = 0.01

    def feature_channels(self, res: int) -> int:
        return ([self.post_filters[-1]] + list(self.up_filters))[res]

    def featuremap_resolution(self, input_res: int, res: int) -> int:
        return (input_res // (2 ** len(self.down_filters))) * (2 ** res)


def _randn(gen: torch.Generator, *shape) -> torch.Tensor:
    return torch.randn(tuple(shape), generator=gen, dtype=torch.float32)


def init_generator_params(cfg: GeneratorConfig = GeneratorConfig(), seed: int = 0,
                          perturb: float = 0.0) -> Bundle:
    """Random-init generator bundle.

    ``perturb > 0`` adds ``N(0, perturb)`` to every parameter that the reference
    initialises to exactly zero (``noise_strength``, conv/ToRGB ``bias``,
    ``color_bias``) so those terms are exercised.
    """
    g = torch.Generator().manual_seed(seed)
    p: Bundle = {}
    # mapping (networks.py:215-252): weight = randn/lr_mul, bias = 0
    for i in range(cfg.mapping_layers):
        p[f'mapping.fc{i}.weight'] = _randn(g, cfg.w_dim, cfg.z_dim if i == 0 else cfg.w_dim) / cfg.mapping_lr_multiplier
        p[f'mapping.fc{i}.bias'] = torch.zeros(cfg.w_dim)
    p['mapping.w_avg'] = torch.zeros(cfg.w_dim)
    for res in cfg.block_resolutions:
        cout = cfg.channels(res)
        if res == 4:
            p['synthesis.b4.const'] = _randn(g, cout, 4, 4)
        convs = ([('conv0', cfg.block_in_channels(res))] if res > 4 else []) + [('conv1', cout)]
        for name, cin in convs:
            k = f'synthesis.b{res}.{name}'
            p[f'{k}.weight'] = _randn(g, cout, cin, 3, 3)
            p[f'{k}.bias'] = torch.zeros(cout)
            p[f'{k}.noise_const'] = _randn(g, res, res)
            p[f'{k}.noise_strength'] = torch.zeros(())
            p[f'{k}.affine.weight'] = _randn(g, cin, cfg.w_dim)
            p[f'{k}.affine.bias'] = torch.ones(cin)
    res = cfg.img_resolution
    cin = cfg.channels(res)
    k = f'synthesis.b{res}.torgb'
    p[f'{k}.weight'] = _randn(g, cfg.torgb_out_channels, cin, 1, 1)
    p[f'{k}.bias'] = torch.zeros(cfg.torgb_out_channels)
    p[f'{k}.color_bias'] = torch.zeros(9)
    p[f'{k}.affine.weight'] = _randn(g, cin + 9, cfg.w_dim)
    p[f'{k}.affine.bias'] = torch.ones(cin + 9)
    if perturb > 0:
        for name in list(p.keys()):
            if name.endswith(('.bias', '.noise_strength', '.color_bias')) and 'affine' not in name and 'mapping' not in name:
                p[name] = p[name] + perturb * _randn(g, *p[name].shape)
    return p


def init_encoder_params(cfg: EncoderConfig = EncoderConfig(), seed: int = 1, perturb_bn: float = 0.0) -> Bundle:
    """Xavier-normal conv weights, zero conv bias (factory.py:47-50), default
    BatchNorm buffers.  ``perturb_bn`` randomises the BN affine/running stats so
    that eval-BN folding is exercised."""
    g = torch.Generator().manual_seed(seed)
    p: Bundle = {}
    bn = 2 if cfg.bn_after_activation else 1          # index of the BatchNorm inside the stage's nn.Sequential

    def conv(prefix, cin, cout, k, transposed=False):
        std = math.sqrt(2.0 / ((cin + cout) * k * k))
        p[f'{prefix}.0.weight'] = (_randn(g, cin, cout, k, k) if transposed else _randn(g, cout, cin, k, k)) * std
        p[f'{prefix}.0.bias'] = torch.zeros(cout)
        if perturb_bn > 0 and cfg.bn_after_activation:
            p[f'{prefix}.0.bias'] += perturb_bn * _randn(g, cout)      # (in this order the bias does not fold into the BatchNorm shift)
        p[f'{prefix}.{bn}.weight'] = torch.ones(cout)
        p[f'{prefix}.{bn}.bias'] = torch.zeros(cout)
        p[f'{prefix}.{bn}.running_mean'] = torch.zeros(cout)
        p[f'{prefix}.{bn}.running_var'] = torch.ones(cout)
        if perturb_bn > 0:
            p[f'{prefix}.{bn}.weight'] += perturb_bn * _randn(g, cout)
            p[f'{prefix}.{bn}.bias'] += perturb_bn * _randn(g, cout)
            p[f'{prefix}.{bn}.running_mean'] += perturb_bn * _randn(g, cout)
            p[f'{prefix}.{bn}.running_var'] += perturb_bn * torch.rand(cout, generator=g)

    filters = [cfg.pre_filters] + list(cfg.down_filters)
    conv('encoder.model.0.conv', cfg.in_channels, filters[0], 7)
    idx = 1
    for i in range(1, len(filters)):
        conv(f'encoder.model.{idx}.conv', filters[i - 1], filters[i], 3)
        idx += 1
    filters = [filters[-1]] + list(cfg.post_filters)
    for i in range(1, len(filters)):
        conv(f'encoder.model.{idx}.conv', filters[i - 1], filters[i], 3)
        idx += 1
    filters = [cfg.post_filters[-1]] + list(cfg.up_filters)
    for i in range(1, max(cfg.encode_resolutions) + 1):
        if cfg.bn_after_activation:
            conv(f'decoder.model.{i - 1}.conv', filters[i - 1], filters[i], 3, transposed=True)      # ScaleUpV2: ConvTranspose2d [in, out, 3, 3]
        else:
            conv(f'decoder.model.{i - 1}.conv.conv', filters[i - 1], filters[i], 3)
    return p


def bundle_from_module(module: torch.nn.Module, keep=None) -> Bundle:
    """Flat float32 CPU bundle from a (reference) ``torch.nn.Module``'s
    ``state_dict`` -- the drop-in path for pickled generators."""
    out: Bundle = {}
    for k, v in module.state_dict().items():
        if keep is not None and not keep(k):
            continue
        if v.dtype.is_floating_point:
            out[k] = v.detach().to('cpu', torch.float32).contiguous()
    return out


def bundle_digest(bundle: Bundle) -> str:
    """Order-independent sha256 over the raw float32 bytes; fixtures store this so
    a GPU-box test can assert it regenerated the same weights from the seed."""
    h = hashlib.sha256()
    for k in sorted(bundle.keys()):
        h.update(k.encode())
        h.update(np.ascontiguousarray(bundle[k].detach().cpu().numpy().astype(np.float32)).tobytes())
    return h.hexdigest()


def style_z_from_seed(seed: int, z_dim: int = 64) -> torch.Tensor:
    """``GanPaintEngine.random_style`` / ``SeedBrushLibrary.set_style``
    (forger/ui/brush.py:667-670, forger/ui/library.py:222-225): float64 [1, z_dim]."""
    return torch.from_numpy(np.random.RandomState(seed=seed).randn(1, z_dim))


def interpolated_style_z(seed1: int, seed2: int, alpha: float, z_dim: int = 64) -> torch.Tensor:
    """``SeedBrushLibrary.set_interpolated_style`` (forger/ui/library.py:227-234)."""
    return style_z_from_seed(seed1, z_dim) * alpha + style_z_from_seed(seed2, z_dim) * (1 - alpha)
