"""Geometry encoder (``sauto`` autoencoder, encode path only) on B200.

Mirrors ``BaseGeoEncoder.encode`` -> ``AutoEncoder._encode``
(forger/experimental/autoenc/base.py:123-134, simple_autoencoder.py:289-297): preprocess ->
conv7x7(reflect)+BN+LeakyReLU -> 3x conv3x3 stride 2 -> conv3x3 x2 = g0; bilinear x2 -> conv3x3 = g1.
Eval-mode BatchNorm is folded into each conv's weight/bias at plan time (exact: the default layer
order is conv -> BN -> LeakyReLU, simple_autoencoder.py:102-105) and LeakyReLU(0.01) rides in the
conv epilogue, so each SingleConvolution is one launch.

Two execution modes (same switch as the generator):
* ``'fp32'``: ``nbe_conv2d_f32`` per layer (true FP32, CUDA cores); reflection padding / bilinear resize are torch
  data-movement ops here.
* ``'bf16'``: NHWC bf16 buffers that carry their reflect padding explicitly; the 7x7 first layer runs on CUDA cores
  (``nbe_enc_conv7x7_bf16``, K = 49 is too thin for the tensor pipe), every 3x3 layer (stride 1 or 2) is a *valid*
  tcgen05 implicit GEMM (``nbe_conv_tc_bf16_ex``: TMA traversal stride 2, epilogue writing into the interior of the
  next padded buffer), borders are refreshed by ``nbe_reflect_border_nhwc_bf16`` and ``ScaleUp`` is
  ``nbe_bilinear2x_pad_nhwc_bf16``.  ``encode_into`` writes g0 / g1 straight into the generator's concatenated
  NHWC inputs (torch.cat of networks_modified.py:219 disappears).
"""
from __future__ import annotations

from typing import List

import os
import torch

from . import _lib
from .conv2d_resample import conv2d_f32
from .params import Bundle, EncoderConfig

ACT_LRELU = 3


PREPROC_CODE = {None: 0, 'none': 0, 'inverse': 1, '-11inverse': 2}


def _cs(c: int) -> int:
    """Channel stride of an intermediate NHWC buffer: whole 64-channel TMA boxes (padding channels stay zero)."""
    return (c + 63) // 64 * 64


class GeometryEncoder:
    def __init__(self, params: Bundle, cfg: EncoderConfig = EncoderConfig(), device='cuda', mode: str = 'fp32'):
        assert mode in ('fp32', 'bf16')
        self.mode = mode
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
        self._v2 = bool(getattr(cfg, 'bn_after_activation', False))
        self._slopes = []                  # LeakyReLU slope per entry of self._layers
        if self._v2:
            self._build_v2(params, n_enc, strides, pads)
        else:
            for i in range(n_enc):
                self._layers.append(self._fold(params, f'encoder.model.{i}.conv') + (strides[i], pads[i], False))
            for i in range(max(self.res)):
                self._layers.append(self._fold(params, f'decoder.model.{i}.conv.conv') + (1, 1, True))
            self._slopes = [float(cfg.neg_slope)] * len(self._layers)
        self._n_enc = n_enc
        self._ws = {}                      # (batch, H) -> workspace, least recently used first
        self.max_cached_batch_sizes = 32
        self.max_cached_patches = 1024      # ... and only while their batch sizes sum to no more than this
        self._flat_s2 = os.environ.get('NBE_ENC_PER_TAP') is None      # A/B switch: strided layers on the per-tap kernel
        self._flat_min = int(os.environ.get('NBE_ENC_FLAT_MIN', '32'))  # A/B switch: smallest output side of a strided layer on the flat kernel
        if mode == 'bf16':
            if cfg.in_channels != 1 or cfg.pre_filters <= 0 or cfg.pre_filters % 8 or cfg.preproc_type not in PREPROC_CODE:
                raise RuntimeError('GeometryEncoder: the tensor-core path covers the sauto layout (1-channel input, 7x7 pre-layer)')
            with torch.cuda.device(self.device):
                w0 = self._layers[0][0]
                self._w7 = w0.reshape(w0.shape[0], 49).contiguous()
                self._w7q = torch.zeros((w0.shape[0], 64), dtype=torch.bfloat16, device=self.device)
                self._w7q[:, :49] = self._w7.to(torch.bfloat16)
                self._tc7 = w0.shape[0] % 16 == 0 and w0.shape[0] <= 64
                # block-Toeplitz form of the same layer (csrc/enc7x7_toeplitz.cu): 64 output channels, NBE_ENC7_IM2COL keeps the
                # im2col kernel for A/B
                self._toep7 = w0.shape[0] == 64 and os.environ.get('NBE_ENC7_IM2COL') is None
                if self._toep7:
                    self._w7t = torch.empty((7, 512, 16), dtype=torch.bfloat16, device=self.device)
                    _lib.call('nbe_enc_conv7x7_toeplitz_weights', _lib.ptr(self._w7), _lib.ptr(self._w7t), 64, _lib.stream())
                self._wq = [None]
                for (w, b, stride, pad, up) in self._layers[1:]:
                    cout, cin = w.shape[0], w.shape[1]
                    if cout % 16:
                        raise RuntimeError('GeometryEncoder: tensor-core path needs out-channels in multiples of 16')
                    wq = torch.empty((9, cout, _cs(cin)), dtype=torch.bfloat16, device=self.device)
                    _lib.call('nbe_prepare_weights_bf16', _lib.ptr(w), _lib.ptr(wq), cout, cin, 3, 0, _lib.stream())
                    self._wq.append(wq)
                if self._v2:
                    self._dec_wq = []
                    for (w_conv, w_tr, b) in self._dec:
                        cout, cin = w_tr.shape[0], w_tr.shape[1]
                        if cout % 128 or max(self.res) > 1:
                            raise RuntimeError('GeometryEncoder: the tensor-core path of the --neg_slope variant covers one ScaleUpV2 stage with a '
                                               'multiple of 128 output channels')
                        wq = torch.empty((9, cout, _cs(cin)), dtype=torch.bfloat16, device=self.device)
                        _lib.call('nbe_prepare_weights_bf16', _lib.ptr(w_tr), _lib.ptr(wq), cout, cin, 3, 0, _lib.stream())
                        self._dec_wq.append(wq)
                    f = torch.zeros((4, 4), dtype=torch.float32, device=self.device)
                    f[1, 1] = 1.0               # the FIR pass as a crop: y[oy, ox] = T[oy + 1, ox + 1] (ConvTranspose2d padding 1, output_padding 1)
                    self._f_crop = f
                    self._v2_ws = {}

    def _bn_scale_shift(self, p, prefix):
        g, beta = p[f'{prefix}.weight'].double(), p[f'{prefix}.bias'].double()
        m, v = p[f'{prefix}.running_mean'].double(), p[f'{prefix}.running_var'].double()
        scale = g / torch.sqrt(v + self.cfg.bn_eps)
        return scale, beta - m * scale

    def _build_v2(self, p, n_enc, strides, pads):
        """The ``--neg_slope`` variant: every stage is conv -> LeakyReLU -> eval BatchNorm (simple_autoencoder.py:100-103) and
        the decoder stages are ``ScaleUpV2`` (ConvTranspose2d 3x3 stride 2 -> LeakyReLU -> BatchNorm, :128-148).  The BatchNorm
        of stage k, y = s * a + t, is folded FORWARD into the convolution of stage k+1 -- conv(y) = conv_{w * s}(a) + sum_i t_i
        sum_taps w[o, i], exact because a reflect-padded constant field is constant -- so every stage stays one
        conv + bias + LeakyReLU launch; only the feature maps that leave the encoder (g0, and every ScaleUpV2 output) get
        their BatchNorm as an explicit per-channel scale / shift (``nbe_affine_*``), and the transposed convolutions read those."""
        cfg, dev = self.cfg, self.device
        n_pre_down = 1 + len(cfg.down_filters)
        prev = None
        for i in range(n_enc):
            prefix = f'encoder.model.{i}.conv'
            w, b = p[f'{prefix}.0.weight'].double(), p[f'{prefix}.0.bias'].double()
            if prev is not None:
                b = b + (w.sum(dim=(2, 3)) * prev[1][None, :]).sum(dim=1)
                w = w * prev[0][None, :, None, None]
            self._layers.append((w.to(dev, torch.float32).contiguous(), b.to(dev, torch.float32).contiguous(), strides[i], pads[i], False))
            self._slopes.append(float(cfg.neg_slope if i < n_pre_down else cfg.post_neg_slope))
            prev = self._bn_scale_shift(p, f'{prefix}.2')
        self._feat_affine = [tuple(t.to(dev, torch.float32).contiguous() for t in prev)]      # BatchNorm of g0, then of every ScaleUpV2
        self._dec = []                      # ScaleUpV2 stages: (conv-form weight [Cout,Cin,3,3], ConvTranspose2d weight as [Cout,Cin,3,3], bias)
        for i in range(max(self.res)):
            prefix = f'decoder.model.{i}.conv'
            wt = p[f'{prefix}.0.weight'].to(dev, torch.float32)                       # [Cin, Cout, 3, 3]
            w_tr = wt.permute(1, 0, 2, 3).contiguous()                                 # T[2Y+kh, 2X+kw] += w_tr[o, i, kh, kw] x[Y, X]
            w_conv = w_tr.flip([2, 3]).contiguous()                                    # the same as a correlation over the zero-stuffed input
            self._dec.append((w_conv, w_tr, p[f'{prefix}.0.bias'].to(dev, torch.float32).contiguous()))
            self._feat_affine.append(tuple(t.to(dev, torch.float32).contiguous() for t in self._bn_scale_shift(p, f'{prefix}.2')))

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

    # ---- tensor-core path ----------------------------------------------------------------------------
    def _workspace(self, B: int, H: int):
        key = (B, H)
        ws = self._ws.get(key)
        if ws is None:
            ws = []
            h = H
            for i, (w, b, stride, pad, up) in enumerate(self._layers):
                if up:
                    ws.append(torch.zeros((B, 2 * h + 2, 2 * h + 2, _cs(w.shape[1])), dtype=torch.bfloat16, device=self.device))
                    h *= 2
                else:
                    h //= stride
                # output of layer i, padded for the next conv
                ws.append(torch.zeros((B, h + 2, h + 2, _cs(w.shape[0])), dtype=torch.bfloat16, device=self.device))
            # last entry: the reflect-padded bf16 copy of the input the block-Toeplitz first layer reads ([B, H+6, W+16])
            ws.append(torch.empty((B, H + 6, H + 16), dtype=torch.bfloat16, device=self.device))
            self._ws[key] = ws
            while len(self._ws) > 1 and (len(self._ws) > self.max_cached_batch_sizes or sum(k[0] for k in self._ws) > self.max_cached_patches):     # least recently used first; a CUDA-graph session that
                self._ws.pop(next(iter(self._ws)))                 # captured an evicted workspace keeps its own reference
        else:
            self._ws[key] = self._ws.pop(key)
        return ws

    def _launch_conv(self, i, cur, B, h, ho, y_ptr, y_cs, rp, ip, next_scale, st):
        """3x3 layer ``i`` (folded weights + bias + LeakyReLU) from the bordered NHWC buffer ``cur`` ([B, h+2, h+2, cs]) to the
        NHWC destination (y_ptr, channel stride y_cs, row / image pitches rp / ip in pixels), on the kernel that suits its shape."""
        w, b, stride, pad, up = self._layers[i]
        cout, cin = w.shape[0], w.shape[1]
        slope = self._slopes[i]
        if stride == 2 and not up and self._flat_s2 and cin % 64 == 0 and cur.shape[3] == cin and cout % 128 == 0 and h % 2 == 0 and ho >= self._flat_min:
            # down-sampling layer at its algorithmic cost: parity planes of the bordered input on the flat CTA-pair kernel
            # (below 32^2 the per-image tiling pads too much; the per-tap kernel batches images into one tile)
            _lib.call('nbe_conv3x3s2_flat_bf16', _lib.ptr(cur), _lib.ptr(self._wq[i]), y_ptr, B, h, h, cin, cout, y_cs, rp, ip,
                      _lib.ptr(b), slope, 1.0, -1.0, _lib.ptr(next_scale), st)
        elif stride == 1 and self._flat_s2 and cout % 128 == 0 and ho >= 32:
            # ScaleUp conv over the bordered bilinear map: 'valid' flat conv, one pass per 128 output channels
            _lib.call('nbe_conv3x3_flat_bf16', _lib.ptr(cur), _lib.ptr(self._wq[i]), y_ptr, B, ho, ho, cin, cur.shape[3],
                      cur.shape[2], 1, cout, y_cs, rp, ip, None, None, 0, 0.0, _lib.ptr(b), slope, 1.0, -1.0,
                      _lib.ptr(next_scale), st)
        else:
            _lib.call('nbe_conv_tc_bf16_ex', _lib.ptr(cur), _lib.ptr(self._wq[i]), y_ptr, B, ho, ho, cur.shape[3], cur.shape[3],
                      cout, y_cs, 3, 1, stride, cur.shape[1], cur.shape[2], rp, ip, None, None, 0, 0.0, _lib.ptr(b), slope, 1.0,
                      -1.0, _lib.ptr(next_scale), st)

    def graph_workspaces(self, B: int, H: int):
        """Every cached buffer a CUDA graph captured at (batch B, input H) points into besides ``_workspace(B, H)``: a session keeps
        the returned objects alive, because the caches are least-recently-used and may drop them (the --neg_slope layout's
        zero-gapped g0 / transposed-conv buffers are keyed by the feature-map size)."""
        if not self._v2 or self.mode != 'bf16':
            return None
        h = H
        for (_, _, stride, _, _) in self._layers:
            h //= stride
        return self._v2_ws.get((B, h))

    @_lib.profiled('encode')
    def encode_into(self, geom, dests, scales=None, scales_ready=None):
        """Run the bf16 encoder and write feature map ``r`` (index into ``self.res``) into ``dests[r] = (tensor, c_off)``:
        an NHWC bf16 tensor [B, R, pitch >= R, cs] whose channels [c_off, c_off + C_r) of the first R columns receive the
        features (e.g. the generator's zero-gapped concat buffers), optionally multiplied by ``scales[r]`` ([B, C_r] float32,
        the consuming layer's styles).  geom: [B,1,H,W] float32, 0 = stroke.
        ``scales_ready``: CUDA event after which ``scales`` (and the destination buffers) may be touched; the layers before
        the first feature map are issued without waiting for it (the caller computes the styles on another stream)."""
        assert self.mode == 'bf16'
        _lib.require_cuda(geom, 'GeometryEncoder.encode_into')
        geom = geom.to(torch.float32).contiguous()
        B, _, H, W = geom.shape
        assert H == W and (H & (H - 1)) == 0, 'square power-of-two patches'
        res = self.res if isinstance(self.res, (list, tuple)) else [self.res]
        max_res = max(res)
        ws = self._workspace(B, H)
        slope = self._slopes[0]
        n_layers = self._n_enc + (0 if self._v2 else max_res)
        with torch.cuda.device(self.device):
            st = _lib.stream()
            wi = 0
            w, b, _, _, _ = self._layers[0]
            cur = ws[wi]; wi += 1
            if self._toep7 and cur.shape[3] == 64 and H % 8 == 0 and W % 128 == 0:
                nb = _lib.load().nbe_enc_conv7x7_toeplitz_scratch_bytes(B, H, W)
                pad = ws[-1]
                assert pad.numel() * 2 >= nb
                _lib.call('nbe_enc_conv7x7_toeplitz_bf16', _lib.ptr(geom), _lib.ptr(self._w7t), _lib.ptr(b), _lib.ptr(cur), _lib.ptr(pad), nb,
                          B, H, W, 64, 64, slope, PREPROC_CODE[self.cfg.preproc_type], st)
            elif self._tc7:
                _lib.call('nbe_enc_conv7x7_tc_bf16', _lib.ptr(geom), _lib.ptr(self._w7q), _lib.ptr(b), _lib.ptr(cur), B, H, W, w.shape[0],
                          cur.shape[3], slope, PREPROC_CODE[self.cfg.preproc_type], st)
            else:
                _lib.call('nbe_enc_conv7x7_bf16', _lib.ptr(geom), _lib.ptr(self._w7), _lib.ptr(b), _lib.ptr(cur), B, H, W, w.shape[0],
                          cur.shape[3], slope, PREPROC_CODE[self.cfg.preproc_type], st)
            if not (self._toep7 and cur.shape[3] == 64 and H % 8 == 0 and W % 128 == 0):      # (the Toeplitz kernel writes the border itself)
                _lib.call('nbe_reflect_border_nhwc_bf16', _lib.ptr(cur), B, H + 2, W + 2, w.shape[0], cur.shape[3], st)
            h = H
            feat_dense = None          # latest feature map as a dense NHWC tensor [B,h,h,C] (input of the next ScaleUp)
            for i in range(1, n_layers):
                w, b, stride, pad, up = self._layers[i]
                cout = w.shape[0]
                if up:
                    if feat_dense is None:
                        raise RuntimeError('GeometryEncoder: ScaleUp needs the previous feature map (encode resolutions must be 0..k)')
                    upbuf = ws[wi]; wi += 1
                    _lib.call('nbe_bilinear2x_pad_nhwc_bf16', _lib.ptr(feat_dense), _lib.ptr(upbuf), B, h, h, feat_dense.shape[3],
                              feat_dense.shape[3], upbuf.shape[3], st)
                    cur = upbuf
                    h *= 2
                ho = h // stride
                feat_idx = i - (self._n_enc - 1)            # >= 0: this layer's output is feature map `feat_idx`
                nxt = ws[wi]; wi += 1
                is_feat = feat_idx >= 0 and feat_idx in res and not self._v2
                last_layer = i == n_layers - 1
                next_scale = None
                direct = is_feat and last_layer             # nothing else consumes it: write (pre-scaled) straight into the destination
                if self._v2 and last_layer:
                    # --neg_slope variant: the last stage's activation goes to a dense map; its BatchNorm, the hand-over to the
                    # generator and the ScaleUpV2 stage follow in _finish_v2
                    a_last = torch.empty((B, ho, ho, cout), dtype=torch.bfloat16, device=self.device)
                    self._launch_conv(i, cur, B, h, ho, a_last.data_ptr(), cout, ho, ho * ho, None, st)
                    if scales_ready is not None:
                        torch.cuda.current_stream().wait_event(scales_ready)
                        scales_ready = None
                    self._finish_v2(a_last, B, ho, res, dests, scales, st)
                    break
                if is_feat and scales_ready is not None:
                    torch.cuda.current_stream().wait_event(scales_ready)      # first consumer of the styles / destination buffers
                    scales_ready = None
                if direct:
                    dst, c_off = dests[res.index(feat_idx)]
                    assert dst.dtype == torch.bfloat16 and dst.shape[0] == B and dst.shape[1] == ho and dst.shape[2] >= ho \
                        and dst.shape[3] >= c_off + cout
                    y_ptr, y_cs, rp, ip = dst.data_ptr() + 2 * c_off, dst.shape[3], dst.shape[2], ho * dst.shape[2]
                    if scales is not None and scales[res.index(feat_idx)] is not None:
                        next_scale = scales[res.index(feat_idx)].to(self.device, torch.float32).contiguous()
                        assert next_scale.shape == (B, cout)
                    padded_out = None
                elif feat_idx >= 0 and (i + 1 < n_layers and self._layers[i + 1][4]):
                    feat_dense = torch.empty((B, ho, ho, cout), dtype=torch.bfloat16, device=self.device)
                    y_ptr, y_cs, rp, ip = feat_dense.data_ptr(), cout, ho, ho * ho
                    padded_out = None
                else:
                    y_cs, rp, ip = nxt.shape[3], ho + 2, (ho + 2) * (ho + 2)
                    y_ptr = nxt.data_ptr() + 2 * ((ho + 2) + 1) * y_cs
                    padded_out = nxt
                self._launch_conv(i, cur, B, h, ho, y_ptr, y_cs, rp, ip, next_scale, st)
                h = ho
                if padded_out is not None:
                    _lib.call('nbe_reflect_border_nhwc_bf16', _lib.ptr(padded_out), B, h + 2, h + 2, cout, padded_out.shape[3], st)
                    cur = padded_out
                elif is_feat and not direct:
                    # a feature that also feeds the decoder: keep the un-scaled dense copy, hand a (scaled) copy to the destination
                    dst, c_off = dests[res.index(feat_idx)]
                    assert dst.shape[0] == B and dst.shape[1] == ho and dst.shape[2] >= ho and dst.shape[3] >= c_off + cout
                    sc = scales[res.index(feat_idx)] if scales is not None else None
                    if sc is not None:
                        sc = sc.to(self.device, torch.float32).contiguous()
                        assert sc.shape == (B, cout)
                    one, zero = self._identity_affine(cout)
                    # one strided copy (times the consuming layer's styles) instead of a chain of torch element-wise kernels
                    _lib.call('nbe_affine_nhwc_bf16', feat_dense.data_ptr(), cout, ho, ho * ho, dst.data_ptr() + 2 * c_off, dst.shape[3],
                              dst.shape[2], ho * dst.shape[2], B, ho, ho, cout, _lib.ptr(one), _lib.ptr(zero), _lib.ptr(sc), st)
        return dests

    def _identity_affine(self, C):
        cache = self.__dict__.setdefault('_id_affine', {})
        if C not in cache:
            cache[C] = (torch.ones(C, dtype=torch.float32, device=self.device), torch.zeros(C, dtype=torch.float32, device=self.device))
        return cache[C]

    def _affine_f32(self, x, affine):
        y = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.call('nbe_affine_nchw_f32', _lib.ptr(x), _lib.ptr(y), x.shape[0], x.shape[1], x.shape[2] * x.shape[3],
                      _lib.ptr(affine[0]), _lib.ptr(affine[1]), _lib.stream())
        return y

    def _encode_f32_v2(self, x, res, max_res):
        """FP32 parity mode of the --neg_slope variant (see ``_build_v2``): folded conv + LeakyReLU per stage, explicit BatchNorm on
        g0 and on every ScaleUpV2 output; ConvTranspose2d(3, stride 2, padding 1, output_padding 1) = zero-stuffing
        (``upfirdn2d`` with a 1x1 filter, up = 2, padding 1) followed by a valid correlation with the flipped kernel."""
        from .upfirdn2d import upfirdn2d
        for i, (w, b, stride, pad, up) in enumerate(self._layers):
            x = x.contiguous()
            xp = torch.empty((x.shape[0], x.shape[1], x.shape[2] + 2 * pad, x.shape[3] + 2 * pad), dtype=torch.float32, device=x.device)
            with torch.cuda.device(x.device):
                _lib.call('nbe_reflect_pad_nchw_f32', _lib.ptr(x), _lib.ptr(xp), x.shape[0] * x.shape[1], x.shape[2], x.shape[3], pad, 0,
                          _lib.stream())
            x = conv2d_f32(xp, w, padding=0, stride=stride, bias=b, act=ACT_LRELU, alpha=self._slopes[i], gain=1.0)
        x = self._affine_f32(x, self._feat_affine[0])
        results = [x]
        one = torch.ones((1, 1), dtype=torch.float32, device=x.device)
        for j in range(max_res):
            w_conv, _, b = self._dec[j]
            u = upfirdn2d(x, one, up=2, padding=[1, 1, 1, 1])                   # [B, C, 2h + 2, 2h + 2]: x at odd positions, zeros elsewhere
            a = conv2d_f32(u, w_conv, padding=0, stride=1, bias=b, act=ACT_LRELU, alpha=float(self.cfg.neg_slope), gain=1.0)
            x = self._affine_f32(a, self._feat_affine[j + 1])
            results.append(x)
        if not isinstance(res, (list, tuple)):
            res = [res]
        return [results[r] for r in res]

    def _finish_v2(self, a_last, B, h, res, dests, scales, st):
        """Tail of the --neg_slope variant on the bf16 path: g0 = BN(a_last) into its destination (times the consuming layer's
        styles) and, for feature 1, ScaleUpV2: g0 in the zero-gapped flat layout -> transposed conv on tcgen05
        (``nbe_convT3x3s2_flat_bf16``, full (2h+1)^2 result) -> crop to [1, 2h] x [1, 2h] + bias + LeakyReLU (the FIR pass with a
        one-tap filter) -> BatchNorm (+ styles) into the destination."""
        dev = self.device
        C0 = a_last.shape[3]

        def hand_over(src, src_cs, src_rp, src_ip, R, C, affine, feat):
            dst, c_off = dests[res.index(feat)]
            assert dst.dtype == torch.bfloat16 and dst.shape[0] == B and dst.shape[1] == R and dst.shape[2] >= R and dst.shape[3] >= c_off + C
            sc = None
            if scales is not None and scales[res.index(feat)] is not None:
                sc = scales[res.index(feat)].to(dev, torch.float32).contiguous()
                assert sc.shape == (B, C)
            _lib.call('nbe_affine_nhwc_bf16', src, src_cs, src_rp, src_ip, dst.data_ptr() + 2 * c_off, dst.shape[3], dst.shape[2],
                      R * dst.shape[2], B, R, R, C, _lib.ptr(affine[0]), _lib.ptr(affine[1]), _lib.ptr(sc), st)

        if 0 in res:
            hand_over(a_last.data_ptr(), C0, h, h * h, h, C0, self._feat_affine[0], 0)
        if max(res) >= 1:
            key = (B, h)
            wsv = self._v2_ws.get(key)
            if wsv is None:
                cout = self._dec[0][1].shape[0]
                wsv = self._v2_ws[key] = (torch.zeros((B, h, h + 1, _cs(C0)), dtype=torch.bfloat16, device=dev),          # g0, zero-gapped
                                          torch.empty((B, 2 * h + 2, 2 * h + 2, cout), dtype=torch.bfloat16, device=dev),  # T
                                          torch.empty((B, 2 * h, 2 * h, cout), dtype=torch.bfloat16, device=dev))         # lrelu(crop(T) + b)
                while len(self._v2_ws) > 4:
                    self._v2_ws.pop(next(iter(self._v2_ws)))
            xg, T, a1 = wsv
            cout = T.shape[3]
            _lib.call('nbe_affine_nhwc_bf16', a_last.data_ptr(), C0, h, h * h, xg.data_ptr(), xg.shape[3], h + 1, h * (h + 1),
                      B, h, h, C0, _lib.ptr(self._feat_affine[0][0]), _lib.ptr(self._feat_affine[0][1]), None, st)
            _lib.call('nbe_convT3x3s2_flat_bf16', _lib.ptr(xg), _lib.ptr(self._dec_wq[0]), _lib.ptr(T), B, h, h, C0, xg.shape[3], h + 1,
                      cout, cout, 2 * h + 2, (2 * h + 2) * (2 * h + 2), None, st)
            _lib.call('nbe_fir_act_nhwc_bf16', _lib.ptr(T), _lib.ptr(self._f_crop), _lib.ptr(a1), B, 2 * h, 2 * h, cout, 2 * h + 1, 2 * h + 1, 1,
                      cout, 2 * h + 2, (2 * h + 2) * (2 * h + 2), cout, 2 * h, 4 * h * h, 1.0, None, None, 0, 0.0, _lib.ptr(self._dec[0][2]),
                      float(self.cfg.neg_slope), 1.0, -1.0, None, st)
            if 1 in res:
                hand_over(a1.data_ptr(), cout, 2 * h, 4 * h * h, 2 * h, cout, self._feat_affine[1], 1)

    def _encode_bf16(self, geom, res) -> List[torch.Tensor]:
        B, H = geom.shape[0], geom.shape[2]
        rl = self.res if isinstance(self.res, (list, tuple)) else [self.res]
        n_down = len(self.cfg.down_filters)
        dests = []
        for r in rl:
            R = (H // (2 ** n_down)) * (2 ** r)
            dests.append((torch.zeros((B, R, R, self.cfg.feature_channels(r)), dtype=torch.bfloat16, device=self.device), 0))
        self.encode_into(geom, dests)
        outs = []
        for (tns, _), r in zip(dests, rl):
            C, R = tns.shape[3], tns.shape[1]
            o = torch.empty((B, C, R, R), dtype=torch.float32, device=self.device)
            _lib.call('nbe_unpack_nchw_f32', _lib.ptr(tns), _lib.ptr(o), B, C, R, R, C, _lib.stream())
            outs.append(o)
        return outs

    def encode(self, geom, res=None) -> List[torch.Tensor]:
        """geom: [B,1,H,W] float, 0 = stroke, 1 = background -> list of float32 NCHW feature maps."""
        _lib.require_cuda(geom, 'GeometryEncoder.encode')
        if res is None:
            res = self.res
        if self.mode == 'bf16' and res == self.res:
            return self._encode_bf16(geom, res)
        x = self.preprocess(geom.to(torch.float32))
        results = []
        max_res = res if not isinstance(res, (list, tuple)) else max(res)
        if self._v2:
            return self._encode_f32_v2(x, res, max_res)
        for i, (w, b, stride, pad, up) in enumerate(self._layers[: self._n_enc + max_res]):
            # reflect padding (of the bilinear x2 map for ScaleUp layers) in one kernel: no torch compute on this path
            x = x.contiguous()
            S = 2 if up else 1
            xp = torch.empty((x.shape[0], x.shape[1], S * x.shape[2] + 2 * pad, S * x.shape[3] + 2 * pad), dtype=torch.float32, device=x.device)
            with torch.cuda.device(x.device):
                _lib.call('nbe_reflect_pad_nchw_f32', _lib.ptr(x), _lib.ptr(xp), x.shape[0] * x.shape[1], x.shape[2], x.shape[3], pad,
                          int(bool(up)), _lib.stream())
            x = xp
            x = conv2d_f32(x, w, padding=0, stride=stride, bias=b, act=ACT_LRELU, alpha=self.cfg.neg_slope, gain=1.0)
            if i >= self._n_enc - 1:
                results.append(x)
        if not isinstance(res, (list, tuple)):
            res = [res]
        return [results[r] for r in res]
