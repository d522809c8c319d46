"""Geometry encoder (``sauto`` autoencoder, encode path only) on B200.

Mirrors ``BaseGeoEncoder.encode`` -> ``AutoEncoder._encode``
(forger/experimental/autoenc/base.py:123-134, simple_autoencoder.py:289-297): preprocess ->
conv7x7(reflect)+BN+LeakyReLU -> 3x conv3x3 stride 2 -> conv3x3 x2 = g0; bilinear x2 -> conv3x3 = g1.
Eval-mode BatchNorm is folded into each conv's weight/bias at plan time (exact: the default layer
order is conv -> BN -> LeakyReLU, simple_autoencoder.py:102-105) and LeakyReLU(0.01) rides in the
conv epilogue, so each SingleConvolution is one ``nbe_conv2d_f32`` launch.  Reflection padding and
the bilinear x2 resize are pure data movement and stay on torch in this round.
"""
from __future__ import annotations

from typing import List

import torch
import torch.nn.functional as F

from . import _lib
from .conv2d_resample import conv2d_f32
from .params import Bundle, EncoderConfig

ACT_LRELU = 3


class GeometryEncoder:
    def __init__(self, params: Bundle, cfg: EncoderConfig = EncoderConfig(), device='cuda'):
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('GeometryEncoder: a CUDA device is required (no CPU fallback on this path)')
        _lib.load()
        self.res = list(cfg.encode_resolutions)
        self._layers = []     # (weight, bias, stride, pad, upsample_before)
        n_enc = 1 + len(cfg.down_filters) + len(cfg.post_filters)
        strides = [1] + [2] * len(cfg.down_filters) + [1] * len(cfg.post_filters)
        pads = [3] + [1] * (n_enc - 1)
        for i in range(n_enc):
            self._layers.append(self._fold(params, f'encoder.model.{i}.conv') + (strides[i], pads[i], False))
        for i in range(max(self.res)):
            self._layers.append(self._fold(params, f'decoder.model.{i}.conv.conv') + (1, 1, True))
        self._n_enc = n_enc

    def _fold(self, p, prefix):
        """conv -> eval BatchNorm  ==  conv with w' = w * g/sqrt(v+eps), b' = (b - m) * g/sqrt(v+eps) + beta."""
        w = p[f'{prefix}.0.weight'].double()
        b = p[f'{prefix}.0.bias'].double()
        g, beta = p[f'{prefix}.1.weight'].double(), p[f'{prefix}.1.bias'].double()
        m, v = p[f'{prefix}.1.running_mean'].double(), p[f'{prefix}.1.running_var'].double()
        scale = g / torch.sqrt(v + self.cfg.bn_eps)
        w = (w * scale[:, None, None, None]).to(self.device, torch.float32).contiguous()
        b = ((b - m) * scale + beta).to(self.device, torch.float32).contiguous()
        return (w, b)

    # reference API -------------------------------------------------------------------------------
    def feature_channels(self, res=0):
        return self.cfg.feature_channels(res)

    def featuremap_resolution(self, input_res, res=0):
        return self.cfg.featuremap_resolution(input_res, res)

    def set_default_encode_resolutions(self, res):
        self.res = list(res) if isinstance(res, (list, tuple)) else res

    def preprocess(self, x):
        t = self.cfg.preproc_type
        if t in (None, 'none'):
            return x
        if t == '-11inverse':
            return (1 - x) * 2 - 1
        if t == 'inverse':
            return 1 - x
        raise RuntimeError(f'Unknown preprocessing type "{t}"')

    def encode(self, geom, res=None) -> List[torch.Tensor]:
        """geom: [B,1,H,W] float, 0 = stroke, 1 = background -> list of float32 NCHW feature maps."""
        _lib.require_cuda(geom, 'GeometryEncoder.encode')
        if res is None:
            res = self.res
        x = self.preprocess(geom.to(torch.float32))
        results = []
        max_res = res if not isinstance(res, (list, tuple)) else max(res)
        for i, (w, b, stride, pad, up) in enumerate(self._layers[: self._n_enc + max_res]):
            if up:
                x = F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=True)
            x = F.pad(x, (pad, pad, pad, pad), mode='reflect')
            x = conv2d_f32(x, w, padding=0, stride=stride, bias=b, act=ACT_LRELU, alpha=self.cfg.neg_slope, gain=1.0)
            if i >= self._n_enc - 1:
                results.append(x)
        if not isinstance(res, (list, tuple)):
            res = [res]
        return [results[r] for r in res]
