"""NeuBE generator forward on B200: mapping -> modulated-conv synthesis with geometry injection -> triad ToRGB.

Host-side mirror of the reference wrapper
(thirdparty/stylegan2_ada_pytorch/training/networks_modified.py:228-400 ``Generator``,
:28-223 ``SynthesisNetwork``; layers from training/networks.py:303-391, 416-485, 540-680):
same call surface -- ``G(z, c, geom_feature, positions=..., noise_buffers=..., truncation_psi=...,
return_debug_data=..., return_features=..., blended_features=..., **synthesis_kwargs)``,
``G.forward_pre_mapped(ws, geom_feature, ...)``, ``G.mapping(z, c, ...)``, ``G.synthesis(ws, geom_feature, ...)``,
attributes ``z_dim, c_dim, w_dim, img_resolution, img_channels, num_ws`` and the debug dict keys
``'uvs','colors','ws','features{res}','features{res}_preblend'``.

Two execution modes, selected like the reference selects precision (``force_fp32`` kwarg):
* ``'bf16'`` (default, = the reference's mixed-fp16 default): NHWC bf16 activations, every modulated conv
  is one tcgen05 implicit-GEMM launch (``nbe_conv_tc_bf16``) with demod/noise/bias/lrelu/clamp and the next
  layer's modulation fused in its epilogue; up-sampling layers are FIR-first (``nbe_upsample2x_nhwc_bf16``).
* ``'fp32'`` (``force_fp32=True``): NCHW float32 on CUDA cores (``nbe_conv2d_f32``), for <= 1e-4 parity.

Forward only.  All compute is in libnbe_b200.so; torch provides memory, streams and RNG.
"""
from __future__ import annotations

import ctypes
import math
import os
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib
from . import upfirdn2d as _up
from .conv2d_resample import conv2d_f32
from .modconv import demod_coefs, weight_sqsum
from .params import Bundle, GeneratorConfig

_NO_STYLES_IN4 = os.environ.get('NBE_NO_STYLES_IN4') is not None      # A/B switch: b4 input by the torch expression

SQRT2 = math.sqrt(2.0)
ACT_LINEAR, ACT_LRELU, ACT_TANH = 1, 3, 4


def fully_connected(x, weight, bias, activation=ACT_LINEAR, lr_multiplier=1.0, act_gain=1.0, alpha=0.0,
                    normalize=False, out=None):
    """``FullyConnectedLayer.forward`` (networks.py:109-122) on ``nbe_fc_f32``."""
    _lib.require_cuda(x, 'fully_connected')
    assert x.ndim == 2 and x.stride(1) == 1
    N, In = x.shape
    Out = weight.shape[0]
    if out is None:
        out = torch.empty((N, Out), dtype=torch.float32, device=x.device)
    is64 = x.dtype == torch.float64
    assert is64 or x.dtype == torch.float32
    with torch.cuda.device(x.device):
        _lib.call('nbe_fc_f32', _lib.ptr(x), int(is64), _lib.ptr(weight), _lib.ptr(bias), _lib.ptr(out), N, In, Out,
                  x.stride(0), out.stride(0), float(lr_multiplier / math.sqrt(In)), float(lr_multiplier),
                  int(activation), float(alpha), float(act_gain), int(bool(normalize)), _lib.stream())
    return out


class WindowBlend:
    """Feature blending on the flat bf16 path (``nbe_blend_window_nhwc_bf16``): the output of block ``res`` is blended with a
    persistent feature canvas and written back into it, per patch window -- the batched form of
    ``PaintingHelper``'s feature canvas (forger/ui/brush.py:190-242) the wavefront scheduler of ``stylizer`` uses.
    fcanvas: [FH, FW, C] bf16; fmask: [FH, FW] uint8; fyx: [B, 2] int32 window origins; base_alpha: [res, res] float32."""

    def __init__(self, res, fcanvas, fmask, fyx, base_alpha, crop_margin):
        self.res, self.fcanvas, self.fmask, self.fyx, self.base_alpha, self.crop_margin = res, fcanvas, fmask, fyx, base_alpha, crop_margin

    def apply(self, x, x_pitch, C, scale, B):
        assert self.fyx.shape == (B, 2) and self.fyx.dtype == torch.int32 and self.fcanvas.shape[2] == C
        _lib.call('nbe_blend_window_nhwc_bf16', _lib.ptr(x), int(x_pitch), int(x.shape[3]), int(self.res), int(C), _lib.ptr(self.fcanvas),
                  _lib.ptr(self.fmask), int(self.fcanvas.shape[0]), int(self.fcanvas.shape[1]), _lib.ptr(self.fyx),
                  _lib.ptr(self.base_alpha), int(self.crop_margin), _lib.ptr(scale), int(B), _lib.stream())


class InjectedGeometry:
    """Geometry features already resident in the generator's concatenated inputs (flat tensor-core path):
    ``buffers[res]`` is the zero-gapped NHWC bf16 tensor [B, res, res + 1, C_block + C_geo] that feeds ``b{2 res}.conv0``;
    its trailing C_geo channels are written by ``GeometryEncoder.encode_into`` *already multiplied by that layer's styles*
    (``scales``), the block's conv1 epilogue fills the leading channels.  Replaces the ``torch.cat([x, geom_feature[i]])``
    of networks_modified.py:219 and the ``x * styles`` of networks.py:68.  Tied to the ``ws`` it was prepared for."""
    def __init__(self, buffers: Dict[int, torch.Tensor], ws, prep):
        self.buffers, self.ws, self.prep = buffers, ws, prep
        # CUDA event after which the geometry channels are in place, when the encoder runs on another stream: the flat path
        # waits for it in front of the first layer that reads them (the blocks before it overlap the encoder)
        self.ready_event = None


class _Layer:
    """Device-side constants of one SynthesisLayer."""
    __slots__ = ('name', 'res', 'up', 'cin', 'cout', 'w32', 'wq', 'wqT', 'wsq', 'bias', 'noise_const', 'noise_strength',
                 'affine_w', 'affine_b')


class MappingNetwork:
    """``MappingNetwork.forward`` (networks.py:255-290), c_dim == 0."""
    def __init__(self, G: 'Generator'):
        self._G = G
        self.z_dim, self.c_dim, self.w_dim, self.num_ws = G.z_dim, 0, G.w_dim, G.num_ws
        self.num_layers = G.cfg.mapping_layers

    def __call__(self, z, c=None, truncation_psi=1, truncation_cutoff=None, skip_w_avg_update=False, broadcast_view=False):
        """``broadcast_view`` (not a reference argument): return the per-layer broadcast as a stride-0 view instead of the
        reference's materialised ``repeat`` -- for callers that only read ws (the engine's fused step)."""
        G = self._G
        _lib.require_cuda(z, 'mapping')
        assert z.ndim == 2 and z.shape[1] == self.z_dim
        x = z if z.dtype in (torch.float32, torch.float64) else z.to(torch.float32)
        x = x.contiguous()
        with _lib.nvtx_range('input'):                              # normalize_2nd_moment is fused into the first FC
            for i in range(self.num_layers):
                x = fully_connected(x, G._map_w[i], G._map_b[i], ACT_LRELU, G.cfg.mapping_lr_multiplier, SQRT2, 0.2,
                                    normalize=(i == 0))
        with _lib.nvtx_range('broadcast'):
            ws = x.unsqueeze(1).expand(-1, self.num_ws, -1) if (broadcast_view and truncation_cutoff is None) \
                else x.unsqueeze(1).repeat([1, self.num_ws, 1])
        if truncation_psi != 1:
            with _lib.nvtx_range('truncate'):
                if truncation_cutoff is None:
                    ws = G._w_avg.lerp(ws, truncation_psi)
                else:
                    ws[:, :truncation_cutoff] = G._w_avg.lerp(ws[:, :truncation_cutoff], truncation_psi)
        return ws


class SynthesisNetwork:
    """``SynthesisNetwork.forward`` (networks_modified.py:123-223), architecture 'orig', colour format 'triad'."""
    def __init__(self, G: 'Generator'):
        self._G = G
        self.w_dim = G.w_dim
        self.img_resolution = G.img_resolution
        self.img_channels = G.img_channels
        self.block_resolutions = G.cfg.block_resolutions
        self.num_ws = G.num_ws
        self.geom_feature_resolutions = list(G.cfg.geom_feature_resolutions)
        self.geom_feature_channels = list(G.cfg.geom_feature_channels)

    def named_buffers(self):
        for L in self._G._layers:
            yield f'{L.name}.noise_const', L.noise_const

    def __call__(self, ws, geom_feature, pos_encoding=None, return_debug_data=False, return_features=None,
                 blended_features=None, noise_buffers=None, **block_kwargs):
        return self._G._synthesis(ws, geom_feature, pos_encoding=pos_encoding, return_debug_data=return_debug_data,
                                  return_features=return_features, blended_features=blended_features,
                                  noise_buffers=noise_buffers, **block_kwargs)

    forward = __call__


class Generator:
    """Drop-in for the reference ``Generator`` on the forward path (see module docstring)."""

    def __init__(self, params: Bundle, cfg: GeneratorConfig = GeneratorConfig(), device='cuda', mode: str = 'bf16'):
        assert mode in ('bf16', 'fp32')
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('Generator: a CUDA device is required (no CPU fallback on this path)')
        _lib.load()
        self.mode = mode
        self.z_dim, self.c_dim, self.w_dim = cfg.z_dim, 0, cfg.w_dim
        self.img_resolution, self.img_channels = cfg.img_resolution, cfg.img_channels
        self.num_ws = cfg.num_ws
        self._build(params)
        self.mapping = MappingNetwork(self)
        self.synthesis = SynthesisNetwork(self)
        self._flat_ws = {}                 # batch size -> workspace, least recently used first
        self.max_cached_batch_sizes = 32   # workspaces kept (least recently used go first) ...
        self.max_cached_patches = 1024     # ... as long as their batch sizes sum to no more than this (~23 MB per patch)
        self._noise_cache = None
        self._window_blend = None          # WindowBlend of the running _synthesis call (flat path)
        self._split = None                 # ('pre' | 'post', res, tensor) of the running _synthesis call (flat path)
        self._noise_prefetched = None      # one-shot (positions, maps) computed ahead of _synthesis by prefetch_noise
        self.last_up_fir_first = os.environ.get('NBE_LAST_UP_FIR_FIRST') is not None   # A/B switch: 4x-FLOP FIR-first path at 128^2
        self.use_flat = os.environ.get('NBE_GEN_V1') is None      # flat shifted-window kernels + algorithmic-cost up-sampling
        self.probe = None       # optional {layer_name: [(start_event, end_event), ...]} filled by _conv_tc (bench.py roofline)
        # A/B switch (default off): NBE_UP_CHUNK_MB=<n> runs every up-sampling layer chunk by chunk through a transposed-conv
        # buffer of n MB, meant to keep T in L2 between the transposed conv and the FIR pass.  Measured (profiles/r02d_chunk_sweep.md):
        # slower at every size -- 12 MB: generator 5.98 ms, 40 MB: 3.66, 96 MB: 3.13 against 2.80 ms unchunked; the per-launch
        # floor of the persistent kernels (~20 us: resident weights, TMEM, tensor maps) outweighs what L2 residency returns.
        self.up_chunk_bytes = int(float(os.environ.get('NBE_UP_CHUNK_MB', '0')) * 2 ** 20)
        self.small_pertap_max = int(os.environ.get('NBE_SMALL_PERTAP_MAX', '0'))   # A/B switch: conv1 of blocks up to this resolution on the per-tap kernel (images batched into a tile)
        self.use_up_fused = os.environ.get('NBE_UP_FUSED') == '1'      # A/B switch: one fused kernel per up layer (csrc/up_fused.cu) instead of transposed conv + FIR pass
        self.up_fused_min_res = int(os.environ.get('NBE_UP_FUSED_MIN_RES', '8'))   # smallest input resolution that takes the fused kernel
        self._up_scratch_bufs = {}
        self.defer_last_layer = False      # flat path: hand the last layer back as a closure instead of launching it (BatchSession)
        self._deferred_last = None

    # ------------------------------------------------------------------------------------------ plan
    def _build(self, p: Bundle) -> None:
        dev = self.device
        f32 = lambda t: t.detach().to(dev, torch.float32).contiguous()
        cfg = self.cfg
        self._map_w = [f32(p[f'mapping.fc{i}.weight']) for i in range(cfg.mapping_layers)]
        self._map_b = [f32(p[f'mapping.fc{i}.bias']) for i in range(cfg.mapping_layers)]
        self._w_avg = f32(p['mapping.w_avg'])
        self._filter = _up.setup_filter([1, 3, 3, 1], device=dev)                 # networks.py:350
        self._const = f32(p['synthesis.b4.const'])                               # [C,4,4]
        self._const_nhwc = self._const.permute(1, 2, 0).contiguous()              # [4,4,C]
        self._lin = {}
        self._layers: List[_Layer] = []
        with torch.cuda.device(dev):
            for res in cfg.block_resolutions:
                for conv in (['conv0'] if res > 4 else []) + ['conv1']:
                    k = f'synthesis.b{res}.{conv}'
                    L = _Layer()
                    L.name, L.res, L.up = f'b{res}.{conv}', res, (2 if conv == 'conv0' else 1)
                    L.w32 = f32(p[f'{k}.weight'])
                    L.cout, L.cin = L.w32.shape[0], L.w32.shape[1]
                    L.wsq = weight_sqsum(L.w32)
                    cin_pad = (L.cin + 63) // 64 * 64
                    L.wq = torch.empty((9, L.cout, cin_pad), dtype=torch.bfloat16, device=dev)
                    # up layers are true convolutions (flip_weight = (up == 1), networks.py:384): pre-flip
                    _lib.call('nbe_prepare_weights_bf16', _lib.ptr(L.w32), _lib.ptr(L.wq), L.cout, L.cin, 3,
                              int(L.up == 2), _lib.stream())
                    L.wqT = None
                    if L.up == 2:                                                  # transposed-conv form: taps as stored (no flip)
                        L.wqT = torch.empty((9, L.cout, cin_pad), dtype=torch.bfloat16, device=dev)
                        _lib.call('nbe_prepare_weights_bf16', _lib.ptr(L.w32), _lib.ptr(L.wqT), L.cout, L.cin, 3, 0, _lib.stream())
                    L.bias = f32(p[f'{k}.bias'])
                    L.noise_const = f32(p[f'{k}.noise_const'])
                    L.noise_strength = float(p[f'{k}.noise_strength'])
                    L.affine_w = f32(p[f'{k}.affine.weight'])
                    L.affine_b = f32(p[f'{k}.affine.bias'])
                    self._layers.append(L)
                self._lin[res] = torch.linspace(0, 1, res).to(dev)                # networks.py:295-299 (CPU linspace)
            k = f'synthesis.b{cfg.img_resolution}.torgb'
            self._rgb_w = f32(p[f'{k}.weight']).reshape(cfg.torgb_out_channels, -1)      # 3 (triad) or 3 + 5 (canvas) rows
            self._canvas_format = cfg.color_format == 'canvas'
            self._rgb_extra = {}
            self._rgb_b = f32(p[f'{k}.bias'])
            self._rgb_color_bias = f32(p[f'{k}.color_bias'])
            self._rgb_affine_w = f32(p[f'{k}.affine.weight'])
            self._rgb_affine_b = f32(p[f'{k}.affine.bias'])
        self._layer_by_name = {L.name: L for L in self._layers}
        # static halves of the fused styles/demod and shifted-noise launch tables (host arrays of device pointers)
        # one table entry = one row range of an affine layer, written as its own contiguous [B, rows] tensor:
        #   every layer's styles (+ demodulation coefficients), the ToRGB affine as its two halves -- 9 colour outputs and the C
        #   styles / sqrt(C) (networks.py:455-462) --, and the block-channel (':head') / injected-geometry-channel (':geo') halves of
        #   the layers that read a concat buffer: nothing is sliced + copied by torch kernels in a forward pass
        wd = self.w_dim
        nl = len(self._layers)
        ent = [dict(key=L.name, aw=L.affine_w.data_ptr(), ab=L.affine_b.data_ptr(), wsq=L.wsq.data_ptr(), cin=L.cin, cout=L.cout,
                    widx=li, pscale=1.0) for li, L in enumerate(self._layers)]   # layer l reads ws[:, l] (networks_modified.py:144-151)
        rows_rgb = self._rgb_affine_w.shape[0]
        rgb_aw, rgb_ab = self._rgb_affine_w.data_ptr(), self._rgb_affine_b.data_ptr()
        ent.append(dict(key='rgb:colors', aw=rgb_aw, ab=rgb_ab, wsq=None, cin=9, cout=0, widx=nl, pscale=1.0))
        ent.append(dict(key='rgb:styles', aw=rgb_aw + 4 * 9 * wd, ab=rgb_ab + 4 * 9, wsq=None, cin=rows_rgb - 9, cout=0, widx=nl,
                        pscale=1.0 / math.sqrt(self._rgb_w.shape[1])))
        self._virt = []
        for li, L in enumerate(self._layers):
            if L.name.endswith('.conv0') and (L.res // 2) in cfg.geom_feature_resolutions:
                cb = cfg.channels(L.res // 2)
                self._virt += [(f'{L.name}:head', li, 0, cb), (f'{L.name}:geo', li, cb, L.cin - cb)]
        if len(ent) + len(self._virt) > 24:                         # (the launch table of nbe_styles_demod_f32 holds 24 entries)
            self._virt = []
        for key, li, r0, rows in self._virt:
            L = self._layers[li]
            ent.append(dict(key=key, aw=L.affine_w.data_ptr() + 4 * r0 * wd, ab=L.affine_b.data_ptr() + 4 * r0, wsq=None, cin=rows,
                            cout=0, widx=li, pscale=1.0))
        ne = len(ent)
        VP, IA, FA = ctypes.c_void_p * ne, ctypes.c_int * ne, ctypes.c_float * ne
        self._tab_ent = ent
        self._tab_n = ne
        self._tab_aw = VP(*[e['aw'] for e in ent])
        self._tab_ab = VP(*[e['ab'] for e in ent])
        self._tab_wsq = VP(*[e['wsq'] for e in ent])
        self._tab_cin = IA(*[e['cin'] for e in ent])
        self._tab_cout = IA(*[e['cout'] for e in ent])
        self._tab_widx = IA(*[e['widx'] for e in ent])
        self._tab_widx0 = IA(*([0] * ne))                           # a single w repeated for every layer: ws passed as [B, 1, w]
        self._tab_pscale = FA(*[e['pscale'] for e in ent])
        self._tab_pfrom = IA(*([0] * ne))
        self._tab_in4_layer = [e['key'] for e in ent].index('b4.conv1')
        self._in4_filled = None                                     # (styles of b4.conv1, in4 tensor) of the last _styles(ws, in4) launch
        self._VP = VP
        nn = len(self._layers)
        self._ntab_nc = (ctypes.c_void_p * nn)(*[L.noise_const.data_ptr() for L in self._layers])
        self._ntab_lin = (ctypes.c_void_p * nn)(*[self._lin[L.res].data_ptr() for L in self._layers])
        self._ntab_res = (ctypes.c_int * nn)(*[L.res for L in self._layers])

    def _workspace(self, B: int):
        """Zero-gapped NHWC bf16 buffers of the flat path, allocated (and zeroed) once per batch size: kernels only ever
        write the valid columns, so the gap columns stay zero across calls."""
        ws = self._flat_ws.get(B)
        if ws is None:
            cfg, dev, bf = self.cfg, self.device, torch.bfloat16
            ws = {}
            for res in cfg.block_resolutions:
                if res < cfg.img_resolution:
                    ctot = cfg.block_in_channels(res * 2)
                    ws[f'out{res}'] = torch.zeros((B, res, res + 1, ctot), dtype=bf, device=dev)       # conv1 output = next conv0 input
                if res > 4:
                    # transposed-conv output T of an up-sampling layer, for ONE chunk of the batch: the layer runs chunk by chunk
                    # (transposed conv -> FIR pass) through this one buffer, sized to stay resident in L2 between the two
                    # kernels, so T is neither written to nor re-read from HBM (see _up_chunk)
                    ws[f't{res}'] = torch.zeros((self._up_chunk(B, res), res + 2, res + 2, cfg.channels(res)), dtype=bf, device=dev)
                    pitch = res if res == cfg.img_resolution else res + 1
                    ws[f'x{res}'] = torch.zeros((B, res, pitch, cfg.channels(res)), dtype=bf, device=dev)        # conv0 output = conv1 input
            ws['in4'] = torch.zeros((B, 4, 5, cfg.channels(4)), dtype=bf, device=dev)
            self._flat_ws[B] = ws
            while len(self._flat_ws) > 1 and (len(self._flat_ws) > self.max_cached_batch_sizes or sum(self._flat_ws) > self.max_cached_patches):    # least recently used first; holders of an evicted
                self._flat_ws.pop(next(iter(self._flat_ws)))           # workspace (CUDA-graph sessions) keep it alive themselves
        else:
            self._flat_ws[B] = self._flat_ws.pop(B)                    # most recently used last
        return ws

    def expand_ws(self, ws: torch.Tensor) -> torch.Tensor:
        """w+ latents as [B, num_ws, w_dim]: brush libraries store projections as [1, num_ws, w] or [1, 1, w]
        (forger/ui/library.py:146-186, "num_ws or 1"); a single w is repeated for every layer, anything else is refused
        before a kernel could read past the tensor."""
        if ws.ndim == 2:
            ws = ws.unsqueeze(1)
        if ws.ndim != 3 or ws.shape[2] != self.w_dim or ws.shape[1] not in (1, self.num_ws):
            raise RuntimeError(f'w+ latents must be [B, {self.num_ws} or 1, {self.w_dim}], got {tuple(ws.shape)}')
        return ws.expand(-1, self.num_ws, -1) if ws.shape[1] == 1 else ws

    def _up_scratch(self, W: int) -> torch.Tensor:
        """Per-CTA rings of transposed-conv rows for ``nbe_up_layer_fused_bf16`` at input width ``W`` (zeroed once: the guard
        columns must stay zero, everything else is rewritten by every launch)."""
        t = self._up_scratch_bufs.get(W)
        if t is None:
            nbytes = int(_lib.load().nbe_up_layer_fused_scratch_bytes(int(W)))
            t = self._up_scratch_bufs[W] = torch.zeros((nbytes,), dtype=torch.uint8, device=self.device)
        return t

    def _up_chunk(self, B: int, res: int) -> int:
        """Images per chunk of an up-sampling layer with output resolution ``res``: as many as keep the chunk's T tensor
        ((res+2)^2 x C bf16 per image) under ``up_chunk_bytes`` (0 = the whole batch in one go)."""
        if self.up_chunk_bytes <= 0:
            return B
        per_img = (res + 2) * (res + 2) * self.cfg.channels(res) * 2
        return max(1, min(B, self.up_chunk_bytes // per_img))

    def alloc_injection(self, ws_latents: torch.Tensor):
        """Prepare the flat path for ``ws_latents`` [B, num_ws, w_dim]: computes all styles, and returns
        ``(InjectedGeometry, dests, scales)`` where ``dests[i] = (tensor, c_off)`` / ``scales[i]`` ([B, C_geo] float32) are
        what ``GeometryEncoder.encode_into`` needs to write feature map i, pre-modulated, into its concat buffer."""
        cfg = self.cfg
        B = ws_latents.shape[0]
        if tuple(ws_latents.shape[1:]) != (self.num_ws, self.w_dim):
            raise RuntimeError(f'alloc_injection: ws must be [B, {self.num_ws}, {self.w_dim}] (see expand_ws), got {tuple(ws_latents.shape)}')
        with torch.cuda.device(self.device):
            wsb = self._workspace(B)
            prep = self._styles(ws_latents.to(torch.float32), in4=wsb['in4'])
        bufs, dests, scales = {}, [], []
        for res, cg in zip(cfg.geom_feature_resolutions, cfg.geom_feature_channels):
            cb = cfg.channels(res)
            t = wsb[f'out{res}']
            bufs[res] = t
            dests.append((t, cb))
            geo = prep[0].get(f'b{res * 2}.conv0:geo')
            scales.append(geo if geo is not None else prep[0][f'b{res * 2}.conv0'][:, cb:cb + cg].contiguous())
        return InjectedGeometry(bufs, ws_latents, prep), dests, scales

    @property
    def flat_supported(self) -> bool:
        """The flat tensor-core path covers the stock configuration (128^2 output, 128 channels in every block); other
        configurations run on the per-tap / generic kernels."""
        cfg = self.cfg
        return self.mode == 'bf16' and self.use_flat and cfg.img_resolution == 128 and \
            all(cfg.channels(r) == 128 for r in cfg.block_resolutions)

    # ------------------------------------------------------------------------------------------ API
    def __call__(self, *args, **kwargs):
        return self.forward(*args, **kwargs)

    def forward(self, z, c=None, geom_feature=None, positions=None, noise_buffers=None, truncation_psi=1,
                truncation_cutoff=None, return_debug_data=False, return_features=None, blended_features=None,
                style_mixing_prob=0, **synthesis_kwargs):
        """networks_modified.py:367-400 (style mixing is training-only and rejected)."""
        if style_mixing_prob and style_mixing_prob > 0:
            raise RuntimeError('Generator.forward: style mixing is a training-time feature (out of scope)')
        ws = self.mapping(z, c, truncation_psi=truncation_psi, truncation_cutoff=truncation_cutoff)
        return self.forward_pre_mapped(ws, geom_feature, positions=positions, return_debug_data=return_debug_data,
                                       return_features=return_features, blended_features=blended_features,
                                       noise_buffers=noise_buffers, **synthesis_kwargs)

    def forward_pre_mapped(self, ws, geom_feature, positions=None, return_debug_data=False, return_features=None,
                           blended_features=None, noise_buffers=None, **synthesis_kwargs):
        """networks_modified.py:346-365."""
        res = self._synthesis(ws, geom_feature, return_debug_data=return_debug_data, return_features=return_features,
                              blended_features=blended_features, noise_buffers=noise_buffers, positions=positions,
                              **synthesis_kwargs)
        if return_debug_data or return_features:
            img, debug = res
            if return_debug_data:
                debug['ws'] = ws
            return img, debug
        return res

    # ------------------------------------------------------------------------------------------ synthesis
    def _noise_for(self, L: _Layer, B: int, noise_mode: str, positions, norm_noise_positions, input_noise):
        """-> (noise tensor or None, batch stride in elements, gain); networks.py:365-382."""
        if noise_mode == 'none':
            return None, 0, 0.0
        R = L.res
        if noise_mode == 'random':
            n = torch.randn([B, 1, R, R], device=self.device)                    # RNG stays in torch (same stream as the reference)
            return n, R * R, L.noise_strength
        assert noise_mode == 'const'
        if positions is not None and input_noise is None and self._noise_cache is not None and self._noise_cache[0] is positions:
            return self._noise_cache[1][L.name], R * R, L.noise_strength
        nc = L.noise_const if input_noise is None else input_noise.to(self.device, torch.float32).contiguous()
        if positions is not None:
            out = torch.empty((B, R, R), dtype=torch.float32, device=self.device)
            pos = positions.to(self.device, torch.int64).contiguous()
            assert pos.shape == (B, 2)
            _lib.call('nbe_shifted_noise_f32', _lib.ptr(nc), _lib.ptr(self._lin[R]), _lib.ptr(pos), _lib.ptr(out),
                      B, R, self.img_resolution, _lib.stream())
            return out, R * R, L.noise_strength
        if norm_noise_positions is not None:
            raise RuntimeError('synthesis: pass integer `positions` (Generator.forward_pre_mapped derives '
                               'norm_noise_positions itself, networks_modified.py:351-353)')
        return nc, 0, L.noise_strength

    def _styles(self, ws: torch.Tensor, in4: Optional[torch.Tensor] = None):
        """All affine layers + demodulation coefficients in one launch (they depend on ws only).  ``in4`` (the flat path's
        zero-gapped b4 input [B, 4, 5, C] bf16): the same launch also writes ``const * styles(b4.conv1)`` into it
        (networks.py:642-643 + :68); ``_run_bf16_flat`` recognises the pair by identity and skips its own torch expression."""
        B = ws.shape[0]
        num_ws, widx = self.num_ws, self._tab_widx
        if ws.shape[1] == self.num_ws and ws.stride(1) == 0 and ws.stride(2) == 1 and ws.stride(0) == self.w_dim:
            num_ws, widx = 1, self._tab_widx0                       # the mapping network's broadcast view: never materialised
        else:
            ws = ws.contiguous()
        dev = self.device
        out_t = [torch.empty((B, e['cin']), dtype=torch.float32, device=dev) for e in self._tab_ent]
        dc_t = [torch.empty((B, e['cout']), dtype=torch.float32, device=dev) if e['cout'] else None for e in self._tab_ent]
        tab_st = self._VP(*[t.data_ptr() for t in out_t])
        tab_dc = self._VP(*[t.data_ptr() if t is not None else None for t in dc_t])
        in_layer = -1
        if _NO_STYLES_IN4:                                          # A/B switch: the torch expression in _run_bf16_flat
            in4 = None
        if in4 is not None:
            in_layer = self._tab_in4_layer
            assert in4.dtype == torch.bfloat16 and in4.is_contiguous() and tuple(in4.shape) == (B,) + tuple(self._const_nhwc.shape[:1]) \
                + (self._const_nhwc.shape[1] + 1, self._const_nhwc.shape[2])
        _lib.call('nbe_styles_demod_input_f32', _lib.ptr(ws), B, num_ws, self.w_dim, self._tab_n, self._tab_aw, self._tab_ab,
                  self._tab_wsq, tab_st, tab_dc, self._tab_cin, self._tab_cout, widx, self._tab_pscale, self._tab_pfrom,
                  in_layer, _lib.ptr(self._const_nhwc) if in4 is not None else None, _lib.ptr(in4), self._const_nhwc.shape[0],
                  self._const_nhwc.shape[1], self._const_nhwc.shape[1] + 1, _lib.stream())
        styles = {e['key']: t for e, t in zip(self._tab_ent, out_t)}
        self._in4_filled = (styles['b4.conv1'], in4) if in4 is not None else None
        dcoefs = {e['key']: t for e, t in zip(self._tab_ent, dc_t) if t is not None}
        # ToRGB: affine -> [colors(9) | styles(C) / sqrt(C)] (networks.py:455-462)
        from .bias_act import bias_act
        colors = bias_act(styles.pop('rgb:colors'), self._rgb_color_bias, dim=1, act='tanh').reshape(-1, 3, 3)
        rgb_styles = styles.pop('rgb:styles')
        return styles, dcoefs, colors, rgb_styles

    def _noise_all(self, B: int, positions):
        """Shifted constant noise of every layer in one launch -> {layer name: [B, R, R]}."""
        pos = positions.to(self.device, torch.int64).contiguous()
        assert pos.shape == (B, 2)
        total = sum(L.res * L.res for L in self._layers)
        buf = torch.empty((B * total,), dtype=torch.float32, device=self.device)
        outs, off = {}, 0
        ptrs = []
        for L in self._layers:
            n = B * L.res * L.res
            outs[L.name] = buf[off:off + n].view(B, L.res, L.res)
            ptrs.append(buf.data_ptr() + 4 * off)
            off += n
        tab_out = (ctypes.c_void_p * len(ptrs))(*ptrs)
        _lib.call('nbe_shifted_noise_all_f32', _lib.ptr(pos), B, self.img_resolution, len(ptrs), self._ntab_nc, self._ntab_lin,
                  tab_out, self._ntab_res, _lib.stream())
        return outs

    def prefetch_noise(self, B: int, positions):
        """Compute the shifted noise maps of the NEXT synthesis call for ``positions`` now (on the current stream); that call
        picks them up if it is given the same ``positions`` tensor object.  One-shot."""
        self._noise_prefetched = (positions, self._noise_all(B, positions)) if positions is not None else None

    def _synthesis(self, ws, geom_feature, pos_encoding=None, return_debug_data=False, return_features=None,
                   blended_features=None, noise_buffers=None, positions=None, noise_mode='random', force_fp32=False,
                   fused_modconv=None, norm_noise_positions=None, window_blend=None, split=None, **unused):
        if pos_encoding is not None:
            raise RuntimeError('synthesis: positional-encoding injection is not part of the style1/2 architecture')
        _lib.require_cuda(ws, 'synthesis')
        assert noise_mode in ('random', 'const', 'none')
        return_features = list(return_features or [])
        blended_features = blended_features or {}
        noise_buffers = noise_buffers or {}
        cfg = self.cfg
        B = ws.shape[0]
        assert ws.shape[1:] == (self.num_ws, self.w_dim)
        ws_in = ws
        ws = ws.to(torch.float32)
        mode = 'fp32' if (force_fp32 or self.mode == 'fp32') else 'bf16'
        injected = isinstance(geom_feature, InjectedGeometry)
        flat = mode == 'bf16' and not return_features and not blended_features and self.flat_supported
        if window_blend is not None and not flat:
            raise RuntimeError('synthesis: window_blend needs the flat bf16 path (no force_fp32 / return_features / blended_features)')
        self._window_blend = window_blend
        # split = ('pre', res, out): run up to conv1 of block `res`, whose un-modulated output goes to `out` ([B, res, res + 1, C] bf16,
        # zero gap column), and stop; ('post', res, xin): start at block 2 * res from `xin`, the (blended, modulated) input of its
        # conv0.  The stylizer's phased feature blending runs every patch's layers before / after the blend point at full batch
        # size and only the blend itself in dependency order (stylizer._stylize_blended_phased).
        if split is not None:
            if not flat or window_blend is not None:
                raise RuntimeError('synthesis: split needs the flat bf16 path without window_blend')
            if split[0] not in ('pre', 'post') or split[1] not in cfg.block_resolutions or split[1] > cfg.img_resolution \
                    or (split[1] == cfg.img_resolution and self._canvas_format) \
                    or any(r > split[1] for r in cfg.geom_feature_resolutions) \
                    or (split[0] == 'pre' and split[1] in cfg.geom_feature_resolutions and not injected):
                raise RuntimeError(f'synthesis: cannot split at block {split[1]} (needs a block at or above '
                                   'every geometry-feature injection, and geometry injected by the encoder when it happens at that block)')
        self._split = split
        if injected and not flat:
            raise RuntimeError('synthesis: InjectedGeometry needs the flat bf16 path (no force_fp32 / return_features / blended_features)')
        with torch.cuda.device(self.device):
            if injected and geom_feature.ws is ws_in:
                styles, dcoefs, colors, rgb_styles = geom_feature.prep
            else:
                if injected:
                    raise RuntimeError('synthesis: InjectedGeometry was prepared for a different ws tensor')
                styles, dcoefs, colors, rgb_styles = self._styles(ws, in4=self._workspace(B)['in4'] if (flat and not (split and split[0] == 'post')) else None)
            run = self._run_fp32 if mode == 'fp32' else (self._run_bf16_flat if flat else self._run_bf16)
            pre, self._noise_prefetched = self._noise_prefetched, None
            if positions is not None and noise_mode == 'const':
                maps = pre[1] if (pre is not None and pre[0] is positions) else self._noise_all(B, positions)
                self._noise_cache = (positions, maps)
            else:
                self._noise_cache = None
            img, uvs, feats = run(B, styles, dcoefs, colors, rgb_styles, geom_feature, positions, norm_noise_positions,
                                  noise_mode, noise_buffers, return_features, blended_features)
        debug = dict(feats)
        if return_debug_data:
            debug.update(self._rgb_extra)                           # 'canvas', 'alpha_fg', 'alpha' in the canvas colour format
            debug['colors'] = colors
            debug['uvs'] = uvs
        if len(debug) > 0:
            return img, debug
        return img

    # ---- FP32 mode ---------------------------------------------------------------------------
    def _run_fp32(self, B, styles, dcoefs, colors, rgb_styles, geom_feature, positions, nnp, noise_mode, noise_buffers,
                  return_features, blended_features):
        cfg = self.cfg
        feats = {}
        x = self._const.unsqueeze(0).expand(B, -1, -1, -1).contiguous()
        geo_idx = 0
        for res in cfg.block_resolutions:
            for conv in (['conv0'] if res > 4 else []) + ['conv1']:
                L = self._layer_by_name[f'b{res}.{conv}']
                noise, nsn, ngain = self._noise_for(L, B, noise_mode, positions, nnp, noise_buffers.get(f'{L.name}.noise_const'))
                if L.up == 2:
                    # FIR-first form of conv2d_resample's up path: pad = 1 + (fw+up-1)//2, 1 + (fw-up)//2 = (3, 2)
                    x = _up.upfirdn2d(x, self._filter, up=2, padding=[3, 2, 3, 2], gain=4)
                    pad, flip = 0, True
                else:
                    pad, flip = 1, False
                x = conv2d_f32(x, L.w32, padding=pad, flip=flip, xscale=styles[L.name], dcoef=dcoefs[L.name],
                               noise=noise, noise_gain=ngain, bias=L.bias, act=ACT_LRELU, alpha=0.2, gain=SQRT2,
                               clamp=cfg.conv_clamp if cfg.conv_clamp is not None else -1)
            if res in return_features:
                feats[f'features{res}_preblend'] = x
            if res in blended_features:
                x = self._blend_nchw(x, blended_features[res])
            if res in return_features:
                feats[f'features{res}'] = x
            if res == cfg.img_resolution:
                img, uvs = self._torgb(x, False, x.shape[1], rgb_styles, colors, B)
            if res in cfg.geom_feature_resolutions:
                x = torch.cat([x, geom_feature[geo_idx].to(self.device, torch.float32)], dim=1)
                geo_idx += 1
        return img, uvs, feats

    def _blend_nchw(self, x, bf):
        saved = bf.features.to(self.device, torch.float32).expand_as(x).contiguous()
        alpha = bf.alpha.to(self.device, torch.float32)
        H, W = x.shape[-2:]
        alpha = alpha.reshape(-1, H, W).contiguous()
        x = x.clone()
        _lib.call('nbe_blend_features', _lib.ptr(x), _lib.ptr(saved), _lib.ptr(alpha), (H * W if alpha.shape[0] > 1 else 0),
                  x.shape[0], x.shape[1], H, W, 0, 0, _lib.stream())
        return x

    def _torgb(self, x, is_bf16, x_cs, rgb_styles, colors, B):
        R = self.img_resolution
        img = torch.empty((B, 3, R, R), dtype=torch.float32, device=self.device)
        uvs = torch.empty((B, 3, R, R), dtype=torch.float32, device=self.device)
        clamp = self.cfg.conv_clamp if self.cfg.conv_clamp is not None else -1
        if self._canvas_format:
            # 'canvas' colour format (networks.py:476-481): generated canvas + 2-way alpha next to the UVS stroke
            canvas = torch.empty((B, 3, R, R), dtype=torch.float32, device=self.device)
            alpha = torch.empty((B, 2, R, R), dtype=torch.float32, device=self.device)
            _lib.call('nbe_torgb_canvas', _lib.ptr(x), int(is_bf16), int(x_cs), _lib.ptr(self._rgb_w), _lib.ptr(rgb_styles),
                      _lib.ptr(self._rgb_b), _lib.ptr(colors.contiguous()), float(clamp), _lib.ptr(img), _lib.ptr(uvs),
                      _lib.ptr(canvas), _lib.ptr(alpha), B, self._rgb_w.shape[1], R, R, _lib.stream())
            self._rgb_extra = {'canvas': canvas, 'alpha_fg': alpha[:, :1], 'alpha': alpha}
            return img, uvs
        _lib.call('nbe_torgb_triad', _lib.ptr(x), int(is_bf16), int(x_cs), _lib.ptr(self._rgb_w), _lib.ptr(rgb_styles),
                  _lib.ptr(self._rgb_b), _lib.ptr(colors.contiguous()), float(clamp), _lib.ptr(img), _lib.ptr(uvs),
                  B, self._rgb_w.shape[1], R, R, _lib.stream())
        return img, uvs

    # ---- BF16 tensor-core mode -----------------------------------------------------------------
    def _conv_tc(self, x, L, B, R, x_cs, y, y_cs, valid, dcoef, noise, nsn, ngain, next_scale):
        clamp = self.cfg.conv_clamp if self.cfg.conv_clamp is not None else -1
        ev = None
        if self.probe is not None and L.name in self.probe:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        _lib.call('nbe_conv_tc_bf16', _lib.ptr(x), _lib.ptr(L.wq), _lib.ptr(y), B, R, R, L.cin, x_cs, L.cout, y_cs, 3,
                  int(valid), _lib.ptr(dcoef), _lib.ptr(noise), nsn, float(ngain), _lib.ptr(L.bias), 0.2, SQRT2,
                  float(clamp), _lib.ptr(next_scale), _lib.stream())
        if ev is not None:
            ev[1].record()
            self.probe[L.name].append(ev)

    def _run_bf16(self, B, styles, dcoefs, colors, rgb_styles, geom_feature, positions, nnp, noise_mode, noise_buffers,
                  return_features, blended_features):
        cfg = self.cfg
        dev = self.device
        bf = torch.bfloat16
        feats = {}
        geo_idx = 0
        x = None          # un-modulated block output, NHWC bf16, channel stride x_cs
        x_cs = 0
        for res in cfg.block_resolutions:
            conv1 = self._layer_by_name[f'b{res}.conv1']
            if res == 4:
                # const input, modulated for b4.conv1 (tiny: B x 4 x 4 x C)
                xin = (self._const_nhwc.unsqueeze(0) * styles[conv1.name][:, None, None, :]).to(bf).contiguous()
            else:
                conv0 = self._layer_by_name[f'b{res}.conv0']
                Rin = res // 2
                U = torch.empty((B, res + 2, res + 2, conv0.cin), dtype=bf, device=dev)
                _lib.call('nbe_upsample2x_nhwc_bf16', _lib.ptr(x), _lib.ptr(self._filter), _lib.ptr(styles[conv0.name]),
                          _lib.ptr(U), B, Rin, Rin, conv0.cin, x_cs, _lib.stream())
                noise, nsn, ngain = self._noise_for(conv0, B, noise_mode, positions, nnp,
                                                    noise_buffers.get(f'{conv0.name}.noise_const'))
                xin = torch.empty((B, res, res, conv0.cout), dtype=bf, device=dev)
                self._conv_tc(U, conv0, B, res, conv0.cin, xin, conv0.cout, True, dcoefs[conv0.name], noise, nsn, ngain,
                              styles[conv1.name])
            # conv1 writes straight into the (possibly concatenated) buffer of the next block's input
            extra = cfg.geom_feature_channels[list(cfg.geom_feature_resolutions).index(res)] \
                if res in cfg.geom_feature_resolutions else 0
            y_cs = conv1.cout + extra
            y = torch.empty((B, res, res, y_cs), dtype=bf, device=dev)
            noise, nsn, ngain = self._noise_for(conv1, B, noise_mode, positions, nnp,
                                                noise_buffers.get(f'{conv1.name}.noise_const'))
            is_last = res == cfg.img_resolution
            fused_rgb = is_last and res not in blended_features and res % 128 == 0 and conv1.cout == 128 and conv1.cin <= 128 \
                and not self._canvas_format
            if fused_rgb:
                # last layer: ToRGB + softmax + colour mix ride in the conv epilogue; the 128-channel map is only stored on request
                need_y = res in return_features
                img = torch.empty((B, 3, res, res), dtype=torch.float32, device=dev)
                uvs = torch.empty((B, 3, res, res), dtype=torch.float32, device=dev)
                clamp = cfg.conv_clamp if cfg.conv_clamp is not None else -1
                ev = None
                if self.probe is not None and conv1.name in self.probe:
                    ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                    ev[0].record()
                _lib.call('nbe_conv_tc_bf16_torgb', _lib.ptr(xin), _lib.ptr(conv1.wq), _lib.ptr(y) if need_y else None, B, res, res,
                          conv1.cin, conv1.cin, conv1.cout, y_cs, 0, _lib.ptr(dcoefs[conv1.name]), _lib.ptr(noise), nsn, float(ngain),
                          _lib.ptr(conv1.bias), 0.2, SQRT2, float(clamp), _lib.ptr(self._rgb_w), _lib.ptr(rgb_styles),
                          _lib.ptr(self._rgb_b), _lib.ptr(colors.contiguous()), float(clamp), _lib.ptr(img), _lib.ptr(uvs),
                          int(need_y), _lib.stream())
                if ev is not None:
                    ev[1].record()
                    self.probe[conv1.name].append(ev)
            else:
                self._conv_tc(xin, conv1, B, res, conv1.cin, y, y_cs, False, dcoefs[conv1.name], noise, nsn, ngain, None)
            x, x_cs = y, y_cs
            if res in return_features:
                feats[f'features{res}_preblend'] = self._unpack(x, B, conv1.cout, res, x_cs)
            if res in blended_features:
                self._blend_nhwc(x, blended_features[res], B, conv1.cout, res, x_cs)
            if res in return_features:
                feats[f'features{res}'] = self._unpack(x, B, conv1.cout, res, x_cs)
            if is_last and not fused_rgb:
                img, uvs = self._torgb(x, True, x_cs, rgb_styles, colors, B)
            if extra:
                g = geom_feature[geo_idx].to(dev, torch.float32).contiguous()
                assert g.shape == (B, extra, res, res), f'geometry feature {tuple(g.shape)} != {(B, extra, res, res)}'
                _lib.call('nbe_pack_nhwc_bf16', _lib.ptr(g), _lib.ptr(x), B, extra, res, res, x_cs, conv1.cout, None,
                          _lib.stream())
                geo_idx += 1
        return img, uvs, feats

    # ---- BF16 tensor-core mode, flat kernels ------------------------------------------------------------
    def _run_bf16_flat(self, B, styles, dcoefs, colors, rgb_styles, geom_feature, positions, nnp, noise_mode, noise_buffers,
                       return_features, blended_features):
        """Every 3x3 layer on the flat shifted-window kernels; up-sampling layers as transposed conv (algorithmic FLOPs) +
        FIR/epilogue; each producer writes its consumer's input already multiplied by the consumer's styles."""
        cfg, dev = self.cfg, self.device
        wsb = self._workspace(B)
        clamp = float(cfg.conv_clamp if cfg.conv_clamp is not None else -1)
        st = _lib.stream()
        injected = isinstance(geom_feature, InjectedGeometry)
        geo_idx = 0
        img = uvs = None
        last = cfg.img_resolution
        split = self._split
        skip_to = split[1] if (split is not None and split[0] == 'post') else 0
        if skip_to == last:
            # split at the output resolution: only ToRGB is left (the last feature map is dense, without a gap column)
            y = split[2]
            assert y.dtype == torch.bfloat16 and y.is_contiguous() and tuple(y.shape[:3]) == (B, last, last)
            img, uvs = self._torgb(y, True, y.shape[3], rgb_styles, colors, B)
            return img, uvs, {}
        if skip_to:
            xin, xin_pitch = split[2], skip_to + 1
            assert xin.dtype == torch.bfloat16 and xin.is_contiguous() and tuple(xin.shape[:3]) == (B, skip_to, skip_to + 1)
        else:
            # b4 input: const * styles(b4.conv1), zero-gapped [B,4,5,C]
            c4 = self._layer_by_name['b4.conv1']
            xin = wsb['in4']
            filled, self._in4_filled = self._in4_filled, None
            if filled is None or filled[0] is not styles[c4.name] or filled[1] is not xin:   # (styles from somewhere else: torch expression)
                xin[:, :, :4, :] = (self._const_nhwc.unsqueeze(0) * styles[c4.name][:, None, None, :]).to(torch.bfloat16)
            xin_pitch = 5
        nvtx = _lib.NVTX
        for res in cfg.block_resolutions:
            if res <= skip_to:
                continue
            conv1 = self._layer_by_name[f'b{res}.conv1']
            if nvtx:
                torch.cuda.nvtx.range_push(f'b{res}: modulated_conv2d + bias_act')
            if res > 4:
                conv0 = self._layer_by_name[f'b{res}.conv0']
                Rin = res // 2
                if injected and geom_feature.ready_event is not None and Rin in cfg.geom_feature_resolutions:
                    torch.cuda.current_stream().wait_event(geom_feature.ready_event)      # first consumer of the geometry channels
                    geom_feature.ready_event = None
                noise, nsn, ngain = self._noise_for(conv0, B, noise_mode, positions, nnp, noise_buffers.get(f'{conv0.name}.noise_const'))
                x1 = wsb[f'x{res}']
                x1_pitch = x1.shape[2]
                if res == last and self.last_up_fir_first:
                    # A/B path (NBE_LAST_UP_FIR_FIRST): FIR-upsample the modulated input first, then the row-resident kernel
                    # in 'valid' mode -- 4x the algorithmic FLOPs; was the faster choice at 128^2 until the FIR pass was tiled
                    U = wsb.get('u128')
                    if U is None:
                        U = wsb['u128'] = torch.empty((B, res + 2, res + 2, conv0.cin), dtype=torch.bfloat16, device=dev)
                    _lib.call('nbe_upsample2x_nhwc_bf16_ex', _lib.ptr(xin), _lib.ptr(self._filter), None, _lib.ptr(U), B, Rin, Rin,
                              conv0.cin, xin.shape[3], xin_pitch, st)
                    _lib.call('nbe_conv_tc_bf16', _lib.ptr(U), _lib.ptr(conv0.wq), _lib.ptr(x1), B, res, res, conv0.cin, conv0.cin,
                              conv0.cout, conv0.cout, 3, 1, _lib.ptr(dcoefs[conv0.name]), _lib.ptr(noise), nsn, float(ngain),
                              _lib.ptr(conv0.bias), 0.2, SQRT2, clamp, _lib.ptr(styles[conv1.name]), st)
                elif self.use_up_fused and Rin >= self.up_fused_min_res and conv0.cout == 128 and conv0.cin <= 128 and Rin % 8 == 0 \
                        and Rin <= 120 and xin_pitch == Rin + 1:
                    # the whole layer in one kernel: transposed conv -> L2-resident ring of T rows -> FIR + epilogue (csrc/up_fused.cu)
                    _lib.call('nbe_up_layer_fused_bf16', _lib.ptr(xin), _lib.ptr(conv0.wqT), _lib.ptr(self._filter), _lib.ptr(x1),
                              _lib.ptr(self._up_scratch(Rin)), self._up_scratch(Rin).numel(), B, Rin, Rin, conv0.cin, xin.shape[3], xin_pitch,
                              conv0.cout, conv0.cout, x1_pitch, res * x1_pitch, 4.0, _lib.ptr(dcoefs[conv0.name]), _lib.ptr(noise), nsn,
                              float(ngain), _lib.ptr(conv0.bias), 0.2, SQRT2, clamp, _lib.ptr(styles[conv1.name]), st)
                else:
                    t = wsb[f't{res}']
                    TP = res + 2
                    Bc = t.shape[0]
                    dco, nsc = dcoefs[conv0.name], styles[conv1.name]
                    for c0 in range(0, B, Bc):
                        n = min(Bc, B - c0)
                        _lib.call('nbe_convT3x3s2_flat_bf16', _lib.ptr(xin[c0:]), _lib.ptr(conv0.wqT), _lib.ptr(t), n, Rin, Rin, conv0.cin,
                                  xin.shape[3], xin_pitch, conv0.cout, conv0.cout, TP, TP * TP, None, st)
                        _lib.call('nbe_fir_act_nhwc_bf16', _lib.ptr(t), _lib.ptr(self._filter), _lib.ptr(x1[c0:]), n, res, res, conv0.cout,
                                  res + 1, res + 1, 1, conv0.cout, TP, TP * TP, conv0.cout, x1_pitch, res * x1_pitch, 4.0,
                                  _lib.ptr(dco[c0:]), _lib.ptr(noise[c0:] if (noise is not None and nsn) else noise), nsn, float(ngain),
                                  _lib.ptr(conv0.bias), 0.2, SQRT2, clamp, _lib.ptr(nsc[c0:]), st)
            else:
                x1, x1_pitch = xin, xin_pitch
            noise, nsn, ngain = self._noise_for(conv1, B, noise_mode, positions, nnp, noise_buffers.get(f'{conv1.name}.noise_const'))
            if res == last and self._canvas_format and self._window_blend is not None and self._window_blend.res == res:
                raise RuntimeError("synthesis: window_blend at the output resolution is not available in the 'canvas' colour format")
            if res == last and self._canvas_format:
                # 'canvas' colour format: 8 ToRGB outputs do not fit the fused epilogue -- store the last feature map, then ToRGB
                y = torch.empty((B, res, res, conv1.cout), dtype=torch.bfloat16, device=dev)
                _lib.call('nbe_conv_tc_bf16', _lib.ptr(x1), _lib.ptr(conv1.wq), _lib.ptr(y), B, res, res, conv1.cin, x1.shape[3],
                          conv1.cout, conv1.cout, 3, 0, _lib.ptr(dcoefs[conv1.name]), _lib.ptr(noise), nsn, float(ngain),
                          _lib.ptr(conv1.bias), 0.2, SQRT2, clamp, None, st)
                img, uvs = self._torgb(y, True, conv1.cout, rgb_styles, colors, B)
                if nvtx:
                    torch.cuda.nvtx.range_pop()
                break
            wb = self._window_blend if (self._window_blend is not None and self._window_blend.res == res) else None
            pre_last = res == last and split is not None and split[0] == 'pre' and split[1] == res
            if res == last and (wb is not None or pre_last):
                # blending at the output resolution: the last feature map must exist in memory, so ToRGB runs as its own kernel
                if pre_last:
                    y = split[2]
                    assert y.dtype == torch.bfloat16 and y.is_contiguous() and tuple(y.shape) == (B, res, res, conv1.cout)
                else:
                    y = torch.empty((B, res, res, conv1.cout), dtype=torch.bfloat16, device=dev)
                _lib.call('nbe_conv_tc_bf16', _lib.ptr(x1), _lib.ptr(conv1.wq), _lib.ptr(y), B, res, res, conv1.cin, x1.shape[3],
                          conv1.cout, conv1.cout, 3, 0, _lib.ptr(dcoefs[conv1.name]), _lib.ptr(noise), nsn, float(ngain),
                          _lib.ptr(conv1.bias), 0.2, SQRT2, clamp, None, st)
                if pre_last:
                    if nvtx:
                        torch.cuda.nvtx.range_pop()
                    return None, None, {'split_next_scale': None}
                wb.apply(y, res, conv1.cout, None, B)
                img, uvs = self._torgb(y, True, conv1.cout, rgb_styles, colors, B)
                if nvtx:
                    torch.cuda.nvtx.range_pop()
                break
            if res == last:
                colors_c = colors.contiguous()

                def run_last(x1=x1, conv1=conv1, noise=noise, nsn=nsn, ngain=ngain, res=res):
                    ev = None
                    if self.probe is not None and conv1.name in self.probe:
                        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                        ev[0].record()
                    img = torch.empty((B, 3, res, res), dtype=torch.float32, device=dev)
                    uvs = torch.empty((B, 3, res, res), dtype=torch.float32, device=dev)
                    with torch.cuda.device(dev):
                        _lib.call('nbe_conv_tc_bf16_torgb', _lib.ptr(x1), _lib.ptr(conv1.wq), None, B, res, res, conv1.cin, conv1.cin,
                                  conv1.cout, conv1.cout, 0, _lib.ptr(dcoefs[conv1.name]), _lib.ptr(noise), nsn, float(ngain),
                                  _lib.ptr(conv1.bias), 0.2, SQRT2, clamp, _lib.ptr(self._rgb_w), _lib.ptr(rgb_styles), _lib.ptr(self._rgb_b),
                                  _lib.ptr(colors_c), clamp, _lib.ptr(img), _lib.ptr(uvs), 0, _lib.stream())
                    if ev is not None:
                        ev[1].record()
                        self.probe[conv1.name].append(ev)
                    return img, uvs

                if self.defer_last_layer:
                    # a CUDA-graph session that wants CUDA events around the dominant kernel captures everything up to here and
                    # launches the last layer (conv + fused ToRGB) eagerly after every replay: see engine.BatchSession
                    self._deferred_last = run_last
                    img = uvs = None
                else:
                    img, uvs = run_last()
                if nvtx:
                    torch.cuda.nvtx.range_pop()
                break
            # conv1 -> next block's (pre-modulated, concatenated, zero-gapped) input
            nxt = self._layer_by_name[f'b{res * 2}.conv0']
            out = geom_feature.buffers[res] if (injected and res in cfg.geom_feature_resolutions) else wsb[f'out{res}']
            if nxt.cin != conv1.cout:
                ns = styles.get(f'{nxt.name}:head')
                if ns is None:
                    ns = styles[nxt.name][:, :conv1.cout].contiguous()
            else:
                ns = styles[nxt.name]
            pre_split = split is not None and split[0] == 'pre' and split[1] == res
            if pre_split:
                # a block whose output buffer also holds injected geometry channels keeps writing there (the encoder's share is
                # already in it) and the whole buffer is copied out afterwards; otherwise conv1 writes straight into the caller's
                copy_out = res in cfg.geom_feature_resolutions
                if not copy_out:
                    out = split[2]
                assert split[2].dtype == torch.bfloat16 and split[2].is_contiguous() and tuple(split[2].shape) == tuple(out.shape) \
                    and out.shape[3] >= conv1.cout
            if res <= self.small_pertap_max and x1_pitch == res + 1 and x1.shape[1] == res:
                # small maps: the flat kernel tiles every image by itself (a 4x4 map fills 20 of a tile's 256 positions); the per-tap
                # kernel packs 128 / res^2 images into one position tile and reads the same zero-gapped layout as a 'same'
                # convolution over a (res) x (res + 1) image whose last column is the zero gap
                _lib.call('nbe_conv_tc_bf16_ex', _lib.ptr(x1), _lib.ptr(conv1.wq), _lib.ptr(out), B, res, res, conv1.cin, x1.shape[3],
                          conv1.cout, out.shape[3], 3, 0, 1, res, res + 1, res + 1, res * (res + 1), _lib.ptr(dcoefs[conv1.name]),
                          _lib.ptr(noise), nsn, float(ngain), _lib.ptr(conv1.bias), 0.2, SQRT2, clamp,
                          _lib.ptr(None if (wb is not None or pre_split) else ns), st)
            else:
                _lib.call('nbe_conv3x3_flat_bf16', _lib.ptr(x1), _lib.ptr(conv1.wq), _lib.ptr(out), B, res, res, conv1.cin, x1.shape[3],
                          x1_pitch, 0, conv1.cout, out.shape[3], res + 1, res * (res + 1), _lib.ptr(dcoefs[conv1.name]), _lib.ptr(noise), nsn,
                          float(ngain), _lib.ptr(conv1.bias), 0.2, SQRT2, clamp, _lib.ptr(None if (wb is not None or pre_split) else ns), st)
            if pre_split:
                if copy_out:
                    if geom_feature.ready_event is not None:
                        torch.cuda.current_stream().wait_event(geom_feature.ready_event)      # the encoder's channels are in place
                        geom_feature.ready_event = None
                    split[2].copy_(out)
                if nvtx:
                    torch.cuda.nvtx.range_pop()
                return None, None, {'split_next_scale': ns}
            if wb is not None:
                wb.apply(out, res + 1, conv1.cout, ns, B)             # blend, save, then the modulation the epilogue would have fused
            if res in cfg.geom_feature_resolutions:
                extra = cfg.geom_feature_channels[list(cfg.geom_feature_resolutions).index(res)]
                if not injected:
                    g = geom_feature[geo_idx].to(dev, torch.float32).contiguous()
                    assert g.shape == (B, extra, res, res), f'geometry feature {tuple(g.shape)} != {(B, extra, res, res)}'
                    tmp = torch.empty((B, res, res, extra), dtype=torch.bfloat16, device=dev)
                    gs = styles[nxt.name][:, conv1.cout:conv1.cout + extra].contiguous()
                    _lib.call('nbe_pack_nhwc_bf16', _lib.ptr(g), _lib.ptr(tmp), B, extra, res, res, extra, 0, _lib.ptr(gs), st)
                    out[:, :, :res, conv1.cout:conv1.cout + extra] = tmp
                geo_idx += 1
            xin, xin_pitch = out, res + 1
            if nvtx:
                torch.cuda.nvtx.range_pop()
        return img, uvs, {}

    def _unpack(self, x, B, C, R, cs):
        out = torch.empty((B, C, R, R), dtype=torch.float32, device=self.device)
        _lib.call('nbe_unpack_nchw_f32', _lib.ptr(x), _lib.ptr(out), B, C, R, R, cs, _lib.stream())
        return out

    def _blend_nhwc(self, x, bfeat, B, C, R, cs):
        saved = bfeat.features.to(self.device, torch.float32).expand(B, C, R, R).contiguous()
        saved_nhwc = torch.empty((B, R, R, cs), dtype=torch.bfloat16, device=self.device)
        _lib.call('nbe_pack_nhwc_bf16', _lib.ptr(saved), _lib.ptr(saved_nhwc), B, C, R, R, cs, 0, None, _lib.stream())
        alpha = bfeat.alpha.to(self.device, torch.float32).reshape(-1, R, R).contiguous()
        _lib.call('nbe_blend_features', _lib.ptr(x), _lib.ptr(saved_nhwc), _lib.ptr(alpha), (R * R if alpha.shape[0] > 1 else 0),
                  B, C, R, R, 1, cs, _lib.stream())
