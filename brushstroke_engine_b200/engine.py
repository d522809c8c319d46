"""Triad paint engine on B200: geometry -> encoder -> generator -> triband alpha/colour composite.

Host-side mirror of ``TriadGanPaintEngine`` / ``GanBrushOptions`` / ``StyleUVSMapper``
(forger/ui/brush.py:410-527, 607-805; forger/ui/mapper.py:16-135):

* ``engine._render_stroke_torch(geom, canvas, opts, **generator_kwargs) -> (result[B,4,W,W], raw, debug_img)``
  with the same argument meaning (geom: float [B,1,W,W], 0 = stroke) and the same ``raw`` dict keys;
* ``engine.render_tiles(...)`` is the batched form of the tail of ``PaintingHelper.render_stroke``
  (brush.py:369-377): crop the margin, x255, clip, truncate to uint8, HWC -- done inside the composite kernel;
* ``GanBrushOptions`` keeps the reference attribute names (style_z, style_ws, color0, color1, canvas_color,
  position, enable_uvs_mapping, custom_args) so caller code reads the same.
"""
from __future__ import annotations

from typing import Dict, Optional

import os

import numpy as np
import torch

from . import _lib
from . import synthetic
from .bias_act import bias_act
from .generator import Generator
from .geo_encoder import GeometryEncoder
from .params import Bundle, EncoderConfig, GeneratorConfig, style_z_from_seed

RENDER_MODES = {'clear': 0, 'full': 1}


class GanBrushOptions:
    """forger/ui/brush.py:410-527."""
    def __init__(self, primary_color=None, secondary_color=None, debug=False):
        self.color0 = primary_color
        self.color1 = secondary_color
        self.canvas_color = None
        self.style_z = None
        self.style_id = None
        self.library_id = ''
        self.style_ws = None
        self.opacity = 1.0
        self.debug = debug
        self.position = None          # [B, 2] int64 (y, x)
        self.custom_args = {}
        self.enable_uvs_mapping = False

    def to(self, device):
        if self.style_z is not None:
            self.style_z = self.style_z.to(device)
        if self.style_ws is not None:
            self.style_ws = self.style_ws.to(device)
        if self.custom_args.get('noise_buffers') is not None:      # WBrushLibrary sets None for plain-tensor styles
            for k, v in self.custom_args['noise_buffers'].items():
                if not torch.is_tensor(v):
                    v = torch.from_numpy(v)
                self.custom_args['noise_buffers'][k] = v.to(device)
        return self

    def set_position(self, x, y):
        if type(x) is int:
            self.position = torch.tensor([y, x], dtype=torch.int64).unsqueeze(0)
        else:
            self.position = torch.stack([y, x], dim=1)

    def get_position(self, device):
        return None if self.position is None else self.position.to(device)

    def set_color(self, color_idx, in_color):
        def prep(x):
            if x is None:
                return None
            color = x if torch.is_tensor(x) else torch.from_numpy(np.asarray(x))
            color = color.to(torch.float32) / 255 if color.dtype == torch.uint8 else color.to(torch.float32)
            return color.unsqueeze(0) if color.ndim == 1 else color
        if color_idx == 0:
            self.color0 = prep(in_color)
        elif color_idx == 1:
            self.color1 = prep(in_color)
        elif color_idx == 2:
            self.canvas_color = prep(in_color)
        else:
            raise ValueError(f'Wrong color idx {color_idx}')

    def set_style(self, style_z, style_id=None):
        self.style_z, self.style_id, self.style_ws = style_z, style_id, None

    def set_style_w(self, style_w, style_id=None, custom_args=None):
        self.style_ws, self.style_id, self.style_z = style_w, style_id, None
        self.custom_args = custom_args if custom_args is not None else {}

    def prepare_style(self, batch_size, device):
        def prep(x):
            if x is None:
                return None
            if x.shape[0] != batch_size:
                assert x.shape[0] == 1, 'Brush options must either have correct style batch, or batch of 1'
                return x.expand(batch_size, *([-1] * (x.ndim - 1))).to(device)
            return x.to(device)
        self.style_z = prep(self.style_z)
        self.style_ws = prep(self.style_ws)

    def prepare_colors(self, default_colors, owned=False):
        """[B,3,ncolors] in [0,1]; user colours override columns (brush.py:514-527).  ``owned``: the caller made
        ``default_colors`` for this call alone, so it is modified in place instead of cloned."""
        out = default_colors if owned else default_colors.clone()
        for idx, col in enumerate((self.color0, self.color1, self.canvas_color)):
            if col is not None:
                out[:, :, idx] = col.to(out.device)
        return out


_ONES3 = {}
_TORCH_COLORS = os.environ.get('NBE_TORCH_COLORS') is not None      # A/B switch: (colors + 1) / 2 as two torch kernels


def _unit_range_colors(colors: torch.Tensor) -> torch.Tensor:
    """``(colors + 1) / 2`` of the generator's tanh colours (brush.py:770,913) as ONE ``nbe_bias_act`` launch -- bias 1, linear,
    gain 0.5: the same float32 add followed by an exact halving -- instead of two torch kernels."""
    if _TORCH_COLORS or colors.dtype != torch.float32 or colors.ndim != 3 or colors.shape[1] != 3 or not colors.is_cuda:
        return (colors + 1) / 2.0
    one = _ONES3.get(colors.device)
    if one is None:
        one = _ONES3[colors.device] = torch.ones(3, dtype=torch.float32, device=colors.device)
    return bias_act(colors.contiguous(), one, dim=1, act='linear', gain=0.5)


class StyleUVSMapper:
    """"Clear background" UVS mapping (forger/ui/mapper.py:16-135).  The reference measures the background S level on
    five bundled 128^2 spline patches; without the reference's image folder the same measurement runs on five
    synthetic cross/stroke patches, or on whatever ``set_geometry`` is given."""
    def __init__(self, engine: 'TriadPaintEngine'):
        self.engine = engine
        self.sfactors: Dict[object, torch.Tensor] = {}
        self.geom_feature = None
        self.bmask = None
        self.fmask = None

    def set_geometry(self, geo_med: torch.Tensor, geo_thick: torch.Tensor):
        """geo_*: [5,1,W,W] float in [0,1], 0 = stroke (medium / thick renderings of the same strokes)."""
        dev = self.engine.device
        geo_med, geo_thick = geo_med.to(dev, torch.float32), geo_thick.to(dev, torch.float32)
        self.geom_feature = self.engine.encoder.encode(geo_med)
        self.fmask = geo_med < 0.01
        self.bmask = geo_thick > 0.99
        self.sfactors = {}

    def _init_geometry(self):
        W = self.engine.patch_width
        med = torch.from_numpy(np.concatenate([synthetic.synthetic_patch(W, seed=10 + i, radius=8) for i in range(5)]))
        thick = torch.from_numpy(np.concatenate([synthetic.synthetic_patch(W, seed=10 + i, radius=12) for i in range(5)]))
        self.set_geometry(med, thick)

    def _render(self, brush_opts, geo_feature):
        G = self.engine.G
        n = geo_feature[0].shape[0]
        if brush_opts.style_ws is not None:
            ws = brush_opts.style_ws.expand(n, -1, -1).to(self.engine.device)
            return G.synthesis(ws, geom_feature=geo_feature, noise_mode='const', return_debug_data=True)
        z = brush_opts.style_z.expand(n, -1).to(self.engine.device)
        return G(z=z, c=None, geom_feature=geo_feature, noise_mode='const', return_debug_data=True)

    def get_colors_raw(self, brush_opts):
        if self.geom_feature is None:
            self._init_geometry()
        _, raw = self._render(brush_opts, [x[:1] for x in self.geom_feature])
        return raw['colors']

    def get_colors(self, brush_opts) -> str:
        """'rgb(r,g,b):rgb(...):rgb(...)' for the brush's three default colours (mapper.py:74-78, 102-103)."""
        colors = ((self.get_colors_raw(brush_opts)[0].detach().cpu() / 2 + 0.5) * 255).to(torch.uint8).numpy()
        return ':'.join('rgb(%s)' % ','.join(str(x) for x in colors[..., i]) for i in range(3))

    def get_brush_icon(self, brush_opts, on_white: bool = True) -> np.ndarray:
        """[W,W,3] uint8 icon: the brush on the first calibration stroke (mapper.py:105-115)."""
        if self.geom_feature is None:
            self._init_geometry()
        renders, raw = self._render(brush_opts, [x[:1] for x in self.geom_feature])
        renders = renders.float()
        if on_white:
            s = raw['uvs'][:, 2:].float()
            renders = renders * (1 - s) + s
        return ((renders[0].permute(1, 2, 0).detach() / 2 + 0.5) * 255).to(torch.uint8).cpu().numpy()

    def get_sfactors(self, opts_list) -> List[torch.Tensor]:
        """``get_sfactor`` for several brushes with ONE generator call (all uncached styles x the 5 calibration patches
        in one batch) -- filling a brush library's cache costs one launch sequence instead of one per brush."""
        if self.geom_feature is None:
            self._init_geometry()
        todo = [o for o in opts_list if o.style_id is None or o.style_id not in self.sfactors]
        fresh = {}
        if todo:
            dev, G = self.engine.device, self.engine.G
            n = self.geom_feature[0].shape[0]
            ws = []
            for o in todo:
                if o.style_ws is not None:
                    w = o.style_ws.to(dev, torch.float32)
                    w = w.expand(-1, G.num_ws, -1) if w.shape[1] == 1 else w
                else:
                    w = G.mapping(o.style_z.to(dev), None).to(torch.float32)
                ws.append(w[:1].expand(n, -1, -1))
            ws = torch.cat(ws).contiguous()
            gf = [x.repeat(len(todo), 1, 1, 1) for x in self.geom_feature]
            _, raw = G.synthesis(ws, geom_feature=gf, noise_mode='const', return_debug_data=True)
            S = raw['uvs'][:, 2:3].reshape(len(todo), n, 1, *raw['uvs'].shape[2:])
            for k, o in enumerate(todo):
                val = torch.stack([torch.topk(S[k, i][self.bmask[i]], k=15)[0].min() for i in range(n)]).min()
                fresh[id(o)] = 1 / val
                if o.style_id is not None:
                    self.sfactors[o.style_id] = fresh[id(o)]
        return [fresh[id(o)] if id(o) in fresh else self.sfactors[o.style_id] for o in opts_list]

    def get_sfactor(self, brush_opts):
        """1 / min_i min(topk15(S_i[background_i])) (mapper.py:117-135); cached per style_id."""
        style_id = brush_opts.style_id
        if style_id is not None and style_id in self.sfactors:
            return self.sfactors[style_id]
        if self.geom_feature is None:
            self._init_geometry()
        _, raw = self._render(brush_opts, self.geom_feature)
        S = raw['uvs'][:, 2:3]
        val = torch.stack([torch.topk(S[i][self.bmask[i]], k=15)[0].min() for i in range(S.shape[0])]).min()
        sfactor = 1 / val
        if style_id is not None:
            self.sfactors[style_id] = sfactor
        return sfactor


class TriadPaintEngine:
    """forger/ui/brush.py:607-805 (``GanPaintEngine`` + ``TriadGanPaintEngine``) built from parameter bundles."""

    def __init__(self, gen_params: Bundle, enc_params: Bundle, device='cuda', mode: str = 'bf16',
                 gen_cfg: GeneratorConfig = GeneratorConfig(), enc_cfg: EncoderConfig = EncoderConfig()):
        self.device = torch.device(device)
        self.G = Generator(gen_params, gen_cfg, self.device, mode=mode)
        self.encoder = GeometryEncoder(enc_params, enc_cfg, self.device, mode=mode)
        self.patch_width = self.G.img_resolution
        self.render_modes = set(RENDER_MODES)
        self.render_mode = 'clear'
        self.style_c = None
        self.uvs_mapper = StyleUVSMapper(self)
        self._side_stream = None
        self._enc_stream = None
        self._overlap_encoder = os.environ.get('NBE_NO_ENCODER_OVERLAP') is None     # A/B switch (graph-captured small batches)
        self._overlap_styles = os.environ.get('NBE_NO_STREAM_OVERLAP') is None
        self._overlap_encoder_all = os.environ.get('NBE_ENC_BRANCH_SMALL_ONLY') is None  # A/B switch: encoder as a graph branch below batch 32 only
        self._batch_sessions = {}
        self._batch_hits = {}
        self._force_overlap = False
        self.use_batch_graph = os.environ.get('NBE_NO_BATCH_GRAPH') is None      # A/B switch: eager launches for the batch step

    def set_render_mode(self, mode):
        if mode not in self.render_modes:
            raise RuntimeError(f'Render mode should be one of {self.render_modes}')
        self.render_mode = mode

    def random_style(self, seed):
        return style_z_from_seed(seed, self.G.z_dim).to(self.device)

    def prepare_geom_input(self, stroke_patch: np.ndarray) -> torch.Tensor:
        """[W,W,C] uint8 (last channel: 255 = stroke) -> [1,1,W,W] float32, 0 = stroke (brush.py:672-681)."""
        g = torch.from_numpy(np.ascontiguousarray(stroke_patch[:, :, -1])).to(self.device)
        return (1 - g.to(torch.float32) / 255.0)[None, None]

    # ------------------------------------------------------------------------------------------------
    def _generate(self, geom, opts, **generator_kwargs):
        opts.to(self.device)
        B = geom.shape[0]
        opts.prepare_style(B, self.device)
        G = self.G
        fused = G.flat_supported and self.encoder.mode == 'bf16' \
            and not generator_kwargs.get('force_fp32', False) and not generator_kwargs.get('return_features') \
            and not generator_kwargs.get('blended_features') \
            and list(self.encoder.res) == list(range(len(G.cfg.geom_feature_resolutions)))
        if fused:
            # mapping first, then the encoder writes g0 / g1 -- already multiplied by the consuming layers' styles --
            # straight into the generator's concatenated, zero-gapped NHWC inputs
            extra = opts.custom_args if opts.style_ws is not None else {}
            positions = opts.get_position(self.device)
            if self._overlap_styles and (B >= 32 or self._force_overlap or torch.cuda.is_current_stream_capturing()):
                # mapping network, per-layer styles / demodulation and the shifted noise maps are small latency-bound kernels
                # that nothing in the encoder needs before its first feature map: they run on a side stream underneath the
                # encoder's first layers (large batches only: at batch 1 the extra stream bookkeeping costs more host time than
                # the overlap saves -- unless the step is being captured into a CUDA graph, where the fork / join becomes two
                # branches of the graph at no host cost).  (Memory allocated there is only ever re-used by later side-stream work, which
                # starts with wait_stream(main), so it cannot be recycled under a main-stream kernel that still reads it.)
                with torch.cuda.device(self.device):
                    main = torch.cuda.current_stream()
                    if self._side_stream is None:
                        self._side_stream = torch.cuda.Stream(device=self.device)
                    side = self._side_stream
                    side.wait_stream(main)
                    with torch.cuda.stream(side):
                        ws = opts.style_ws if opts.style_ws is not None else G.mapping(opts.style_z, self.style_c, broadcast_view=True)
                        ws = G.expand_ws(ws).to(self.device, torch.float32)   # (a single w per patch stays a stride-0 view)
                        inj, dests, scales = G.alloc_injection(ws)
                        if extra.get('noise_buffers') is None:
                            G.prefetch_noise(B, positions)
                        ready = torch.cuda.Event()
                        ready.record(side)
                    if (B < 32 or self._overlap_encoder_all) and torch.cuda.is_current_stream_capturing() and self._overlap_encoder:
                        # inside a CUDA graph the encoder becomes a third branch next to the synthesis blocks that do not read its
                        # features yet (b4 .. b16: seven launches of ~20 us that occupy a few SMs each); the flat path joins it in
                        # front of b32.conv0 (generator.InjectedGeometry.ready_event).  Batch 1: 0.40 -> 0.36 ms per stroke;
                        # batch 256: 4.01 -> 3.92 ms per step (the small launches fill the tails of the encoder's kernels)
                        if self._enc_stream is None:
                            self._enc_stream = torch.cuda.Stream(device=self.device)
                        enc_s = self._enc_stream
                        enc_s.wait_stream(main)
                        with torch.cuda.stream(enc_s):
                            self.encoder.encode_into(geom, dests, scales, scales_ready=ready)
                            inj.ready_event = torch.cuda.Event()
                            inj.ready_event.record(enc_s)
                    else:
                        self.encoder.encode_into(geom, dests, scales, scales_ready=ready)
                    main.wait_event(ready)
            else:
                ws = opts.style_ws if opts.style_ws is not None else G.mapping(opts.style_z, self.style_c, broadcast_view=True)
                ws = G.expand_ws(ws).to(self.device, torch.float32)
                inj, dests, scales = G.alloc_injection(ws)
                self.encoder.encode_into(geom, dests, scales)
            return G.forward_pre_mapped(ws=ws, positions=positions, geom_feature=inj,
                                        return_debug_data=True, noise_mode='const', **extra, **generator_kwargs)
        geom_feature = self.encoder.encode(geom)
        if opts.style_ws is not None:
            return G.forward_pre_mapped(ws=G.expand_ws(opts.style_ws), positions=opts.get_position(self.device),
                                        geom_feature=geom_feature, return_debug_data=True, noise_mode='const',
                                        **opts.custom_args, **generator_kwargs)
        return G(z=opts.style_z, c=self.style_c, positions=opts.get_position(self.device),
                 geom_feature=geom_feature, return_debug_data=True, noise_mode='const', **generator_kwargs)

    def _composite(self, triad_data, opts, B, want_f32=True, crop_margin=None):
        uvs = triad_data['uvs'].contiguous()
        default_colors = _unit_range_colors(triad_data['colors'])
        sfactor = None
        if opts.enable_uvs_mapping:
            sf = getattr(opts, 'sfactor', None)                 # per-patch factors of a multi-session batch (server.StrokeBatcher)
            if sf is None:
                sf = self.uvs_mapper.get_sfactor(opts)
            sfactor = sf.reshape(-1).to(self.device, torch.float32).expand(B).contiguous()
        colors = opts.prepare_colors(default_colors, owned=True).contiguous()
        W = self.patch_width
        out_f32 = torch.empty((B, 4, W, W), dtype=torch.float32, device=self.device) if want_f32 else None
        out_u8 = None
        m = 0
        if crop_margin is not None:
            m = int(crop_margin)
            out_u8 = torch.empty((B, W - 2 * m, W - 2 * m, 4), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.call('nbe_triad_composite', _lib.ptr(uvs), _lib.ptr(colors), _lib.ptr(sfactor), RENDER_MODES[self.render_mode],
                      _lib.ptr(out_f32), _lib.ptr(out_u8), B, W, W, m, _lib.stream())
        return out_f32, out_u8

    def _render_stroke_torch(self, geom, canvas, opts, **generator_kwargs):
        """-> (result [B,4,W,W] float in [0,1] (straight RGBA), raw net output dict, None)."""
        _lib.require_cuda(geom, '_render_stroke_torch')
        result_img, triad_data = self._generate(geom, opts, **generator_kwargs)
        out_f32, _ = self._composite(triad_data, opts, geom.shape[0], want_f32=True)
        return out_f32, triad_data, None

    def render_tiles(self, geom, opts, crop_margin=0, **generator_kwargs):
        """Batched tail of ``PaintingHelper.render_stroke``: -> (uint8 tiles [B,W-2m,W-2m,4] on device, raw)."""
        _lib.require_cuda(geom, 'render_tiles')
        result_img, triad_data = self._generate(geom, opts, **generator_kwargs)
        _, out_u8 = self._composite(triad_data, opts, geom.shape[0], want_f32=False, crop_margin=crop_margin)
        return out_u8, triad_data

    def render_split_pre(self, geom, opts, res: int, out: torch.Tensor) -> torch.Tensor:
        """First half of ``render_tiles`` for the stylizer's phased feature blending: encoder, mapping and the synthesis blocks
        up to ``conv1`` of block ``res``, whose un-modulated bf16 output lands in ``out`` [B, res, res + 1, C] (zero gap column).
        Returns the styles [B, C] the consumer of that feature map is modulated with (applied after the blend)."""
        _lib.require_cuda(geom, 'render_split_pre')
        _, raw = self._generate(geom, opts, split=('pre', int(res), out))
        return raw['split_next_scale']

    def render_split_post(self, xin: torch.Tensor, opts, res: int, crop_margin: int = 0) -> torch.Tensor:
        """Second half: the synthesis blocks above ``res`` from ``xin`` [B, res, res + 1, C] (the blended feature map, already
        multiplied by the next layer's styles), ToRGB and the composite -> uint8 tiles [B, W-2m, W-2m, 4] on the device."""
        _lib.require_cuda(xin, 'render_split_post')
        B, G = xin.shape[0], self.G
        opts.to(self.device)
        opts.prepare_style(B, self.device)
        extra = opts.custom_args if opts.style_ws is not None else {}
        ws = opts.style_ws if opts.style_ws is not None else G.mapping(opts.style_z, self.style_c)
        ws = G.expand_ws(ws).to(self.device, torch.float32).contiguous()
        _, triad_data = G.forward_pre_mapped(ws=ws, positions=opts.get_position(self.device), geom_feature=None, return_debug_data=True,
                                             noise_mode='const', split=('post', int(res), xin), **extra)
        _, out_u8 = self._composite(triad_data, opts, B, want_f32=False, crop_margin=crop_margin)
        return out_u8

    def render_stroke(self, stroke_patch, canvas_patch, opts, **generator_kwargs):
        """[W,W,C] uint8 stroke patch -> ([W,W,4] uint8, None) (brush.py:683-701)."""
        geom = self.prepare_geom_input(stroke_patch)
        tiles, _ = self.render_tiles(geom, opts, crop_margin=0, **generator_kwargs)
        return np.ascontiguousarray(tiles[0].cpu().numpy()), None

    def interactive_session(self, opts: GanBrushOptions, crop_margin: int = 0) -> 'InteractiveSession':
        """CUDA-graph session for one-patch-at-a-time rendering with a fixed brush (see ``InteractiveSession``)."""
        return InteractiveSession(self, opts, crop_margin)

    def batch_session(self, B: int, crop_margin: int = 10) -> 'BatchSession':
        """The CUDA graph of one batch step (encoder -> mapping -> synthesis -> composite -> uint8 tiles) for ``B`` patches
        styled by z, in the engine's current render mode; built on first use and kept per (B, crop_margin, render_mode)."""
        key = (int(B), int(crop_margin), self.render_mode)
        sess = self._batch_sessions.get(key)
        if sess is None:
            sess = self._batch_sessions[key] = BatchSession(self, int(B), int(crop_margin))
            while len(self._batch_sessions) > 4:
                self._batch_sessions.pop(next(iter(self._batch_sessions)))
        return sess

    def render_tiles_graph(self, geom: torch.Tensor, z: torch.Tensor, positions: torch.Tensor, crop_margin: int = 10) -> torch.Tensor:
        """``render_tiles`` for a plain z-styled batch through the batch step's CUDA graph: three small device copies into the
        graph's input buffers + ONE graph launch instead of ~55 kernel launches issued from Python.  Returns the graph's own
        output buffer (valid until the next call with the same batch size; clone it to keep it)."""
        sess = self.batch_session(geom.shape[0], crop_margin)
        return sess.run(geom, z, positions)

    def render_patches_host(self, guidance_patches: torch.Tensor, z: torch.Tensor, positions: torch.Tensor,
                            crop_margin: int = 10, out: Optional[torch.Tensor] = None, wait: bool = True, **generator_kwargs):
        """End-to-end batched entry point with HOST buffers (the stylizer's per-batch work):
        guidance_patches [B,W,W] uint8 (0 = stroke, as sliced from the padded guidance image), z [B,z_dim] float64,
        positions [B,2] int64 (y, x) -- all on the host (pinned for async copies) -> uint8 tiles [B,W-2m,W-2m,4] on the host.

        ``wait=False`` returns ``(out, event)`` instead: the device->host copy runs on a side stream and ``event`` fires
        when ``out`` is complete, so a caller that keeps two pinned output buffers overlaps the download of batch i with
        the compute of batch i+1 (``event.synchronize()`` before reading ``out``)."""
        B, W = guidance_patches.shape[0], self.patch_width
        assert guidance_patches.dtype == torch.uint8 and guidance_patches.shape == (B, W, W)
        dev = self.device
        if self.use_batch_graph and not generator_kwargs and self.G.flat_supported and self.encoder.mode == 'bf16' \
                and z.dtype == torch.float64 and positions.dtype == torch.int64:
            # the whole step as one CUDA-graph launch: the host buffers are copied straight into the graph's inputs
            sess = self.batch_session(B, crop_margin)
            tiles = sess.run_host(guidance_patches, z, positions).clone()     # the graph's output buffer is re-used by the next step
        else:
            d_patches = guidance_patches.to(dev, non_blocking=True)
            d_z = z.to(dev, non_blocking=True)
            d_pos = positions.to(dev, non_blocking=True)
            crops = torch.stack([torch.arange(B, dtype=torch.int32, device=dev) * W, torch.zeros(B, dtype=torch.int32, device=dev)], dim=1).contiguous()
            geom = torch.empty((B, 1, W, W), dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                _lib.call('nbe_gather_geom_patches', _lib.ptr(d_patches), B * W, W, _lib.ptr(crops), _lib.ptr(geom), B, W, _lib.stream())
            opts = GanBrushOptions()
            opts.set_style(d_z)
            opts.position = d_pos
            tiles, _ = self.render_tiles(geom, opts, crop_margin=crop_margin, **generator_kwargs)
        if out is None:
            out = torch.empty(tiles.shape, dtype=torch.uint8, pin_memory=True)
        if wait:
            out.copy_(tiles, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return out
        with torch.cuda.device(dev):
            if getattr(self, '_copy_stream', None) is None:
                self._copy_stream = torch.cuda.Stream(device=dev)
            main = torch.cuda.current_stream()
            self._copy_stream.wait_stream(main)
            with torch.cuda.stream(self._copy_stream):
                out.copy_(tiles, non_blocking=True)
                tiles.record_stream(self._copy_stream)
                done = torch.cuda.Event()
                done.record(self._copy_stream)
        return out, done


class BatchSession:
    """One batch step -- geometry [B,1,W,W] + z [B,z_dim] + canvas positions [B,2] -> uint8 RGBA tiles [B,T,T,4] -- captured
    ONCE into a CUDA graph (mapping / styles / noise as a parallel branch next to the encoder's first layers) and replayed per
    batch.  The stylizer's batches, ``render_patches_host`` and the benchmark step go through it: the host issues three small
    copies and one graph launch per batch instead of ~55 kernel launches, which is what bounds eight processes sharing one
    host (8-GPU end-to-end) and removes the launch gaps between the step's small kernels.  Output bytes are identical to the
    eager ``render_tiles`` (same kernels, same order).  The session owns the workspaces its kernels point into."""

    def __init__(self, engine: 'TriadPaintEngine', B: int, crop_margin: int, split_last_layer: bool = False, blend: Optional[dict] = None):
        """``split_last_layer``: capture everything up to the last synthesis layer and launch that layer (the dominant kernel:
        3x3 modconv @128^2 with ToRGB fused) and the composite eagerly after each replay, so that ``engine.G.probe`` can put
        CUDA events around it inside a timed region (events cannot be timed inside a graph).  Same kernels, same bytes.
        ``blend``: ``dict(res, fcanvas, fmask, base_alpha, crop_margin)`` -- the step blends block ``res`` against that
        persistent feature canvas (``generator.WindowBlend``); the window origins are the extra graph input ``_fyx`` [B,2]
        int32.  The warm-up runs blend against a scratch canvas: only replays touch the real one."""
        self.engine, self.B, self.crop_margin = engine, B, crop_margin
        self.split = bool(split_last_layer)
        self._blend = blend
        self._fyx = torch.zeros((B, 2), dtype=torch.int32, device=engine.device)
        assert not (blend and split_last_layer)
        self._tail = None
        dev, W = engine.device, engine.patch_width
        self.render_mode = engine.render_mode
        self._geom = torch.ones((B, 1, W, W), dtype=torch.float32, device=dev)
        self._z = torch.zeros((B, engine.G.z_dim), dtype=torch.float64, device=dev)
        self._pos = torch.zeros((B, 2), dtype=torch.int64, device=dev)
        self._u8 = torch.full((B, W, W), 255, dtype=torch.uint8, device=dev)           # host-fed variant: uint8 guidance patches
        self._crops = torch.stack([torch.arange(B, dtype=torch.int32, device=dev) * W, torch.zeros(B, dtype=torch.int32, device=dev)], dim=1).contiguous()
        self._stream = torch.cuda.Stream(device=dev)
        # host-fed variant: two staging sets (uint8 patches, z, positions) filled on a copy-in stream + the event that frees each
        self._in_stream = torch.cuda.Stream(device=dev)
        self._in_flip = 0
        self._in_sets = []
        for _ in range(2):
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            self._in_sets.append((torch.empty_like(self._u8), torch.empty_like(self._z), torch.empty_like(self._pos), ev))
        with torch.no_grad(), torch.cuda.device(dev):
            self._stream.wait_stream(torch.cuda.current_stream())
            engine._force_overlap = True                             # warm up the code path the capture takes (side-stream branch)
            try:
                with torch.cuda.stream(self._stream):
                    scratch = None
                    if blend is not None:
                        from .generator import WindowBlend
                        r = blend['res']
                        scratch = WindowBlend(r, torch.zeros((r, r, blend['fcanvas'].shape[2]), dtype=torch.bfloat16, device=dev),
                                              torch.zeros((r, r), dtype=torch.uint8, device=dev), self._fyx, blend['base_alpha'],
                                              blend['crop_margin'])
                    for _ in range(2):                              # warm-up: workspaces, cudaFuncSetAttribute, lazy module loads
                        engine.G._noise_cache = None
                        self._forward(scratch)
                        if self._tail is not None:
                            self._tail()
                self._stream.synchronize()
                engine.G._noise_cache = None
                g = torch.cuda.CUDAGraph()
                n0 = _lib.launch_count()
                real = None
                if blend is not None:
                    real = WindowBlend(blend['res'], blend['fcanvas'], blend['fmask'], self._fyx, blend['base_alpha'], blend['crop_margin'])
                with torch.cuda.graph(g, stream=self._stream):
                    self._out = self._forward(real)
                self.kernels_per_replay = _lib.launch_count() - n0   # libnbe_b200 kernels inside the graph (torch's own copies not counted)
            finally:
                engine._force_overlap = False
            engine.G._noise_cache = None
            self._graph = g
            self._workspaces = (engine.G._flat_ws.get(B), engine.encoder._ws.get((B, W)), engine.encoder.graph_workspaces(B, W))
            torch.cuda.current_stream().wait_stream(self._stream)

    def _forward(self, window_blend=None):
        eng = self.engine
        opts = GanBrushOptions()
        opts.set_style(self._z)
        opts.position = self._pos
        if not self.split:
            kw = {'window_blend': window_blend} if window_blend is not None else {}
            tiles, _ = eng.render_tiles(self._geom, opts, crop_margin=self.crop_margin, **kw)
            return tiles
        G = eng.G
        G.defer_last_layer, G._deferred_last = True, None
        try:
            _, triad = eng._generate(self._geom, opts)
        finally:
            G.defer_last_layer = False
        last = G._deferred_last
        if last is None:
            raise RuntimeError('BatchSession(split_last_layer=True) needs the flat tensor-core path of the stock configuration')
        G._deferred_last = None

        def tail():
            _, uvs = last()
            data = dict(triad)
            data['uvs'] = uvs
            return eng._composite(data, opts, self.B, want_f32=False, crop_margin=self.crop_margin)[1]
        self._tail = tail
        return None

    def run(self, geom: torch.Tensor, z: torch.Tensor, positions: torch.Tensor) -> torch.Tensor:
        """Device-resident inputs -> the graph's output buffer (on the caller's current stream)."""
        self._geom.copy_(geom, non_blocking=True)
        self._z.copy_(z, non_blocking=True)
        self._pos.copy_(positions, non_blocking=True)
        self._graph.replay()
        return self._tail() if self.split else self._out

    def run_host(self, patches_u8: torch.Tensor, z: torch.Tensor, positions: torch.Tensor) -> torch.Tensor:
        """Pinned host inputs (uint8 guidance patches [B,W,W], z float64, positions int64) -> the graph's output buffer.
        The uploads go through two staging sets on a copy-in stream, so the inputs of step i+1 cross PCIe while step i computes
        (the graph's own input buffers are still being read then); the main stream only waits for the upload's event."""
        eng, W = self.engine, self.engine.patch_width
        with torch.cuda.device(eng.device):
            main = torch.cuda.current_stream()
            if self._in_stream is None or os.environ.get('NBE_E2E_SINGLE_STREAM') is not None:      # A/B switch: uploads on the main stream
                u8 = self._u8
                u8.copy_(patches_u8, non_blocking=True)
                self._z.copy_(z, non_blocking=True)
                self._pos.copy_(positions, non_blocking=True)
            else:
                k = self._in_flip
                self._in_flip ^= 1
                u8, zs, ps, free = self._in_sets[k]
                with torch.cuda.stream(self._in_stream):
                    self._in_stream.wait_event(free)                 # the gather of two steps ago is done with this set
                    u8.copy_(patches_u8, non_blocking=True)
                    zs.copy_(z, non_blocking=True)
                    ps.copy_(positions, non_blocking=True)
                    up = torch.cuda.Event()
                    up.record(self._in_stream)
                main.wait_event(up)
                self._z.copy_(zs, non_blocking=True)
                self._pos.copy_(ps, non_blocking=True)
            _lib.call('nbe_gather_geom_patches', _lib.ptr(u8), self.B * W, W, _lib.ptr(self._crops), _lib.ptr(self._geom), self.B, W,
                      _lib.stream())
            if u8 is not self._u8:
                free.record(main)
        self._graph.replay()
        return self._tail() if self.split else self._out


CANVAS_RENDER_MODES = {'clear': 0, 'stroke': 1, 'canvas': 2, 'full': 3}


class CanvasPaintEngine(TriadPaintEngine):
    """forger/ui/brush.py:870-935: the 'canvas' colour format (ToRGB emits a generated canvas and a 2-way alpha next to the
    UVS stroke, networks.py:476-481).  Render modes: 'clear' (stroke colour with the generated foreground alpha), 'stroke'
    (opaque stroke colour), 'canvas' (the generated canvas only), 'full' (canvas under the stroke).  No UVS remapping."""

    def __init__(self, gen_params: Bundle, enc_params: Bundle, device='cuda', mode: str = 'bf16',
                 gen_cfg: Optional[GeneratorConfig] = None, enc_cfg: EncoderConfig = EncoderConfig()):
        gen_cfg = gen_cfg if gen_cfg is not None else GeneratorConfig(color_format='canvas')
        if gen_cfg.color_format != 'canvas':
            raise RuntimeError("CanvasPaintEngine needs a generator with color_format == 'canvas'")
        super().__init__(gen_params, enc_params, device, mode, gen_cfg, enc_cfg)
        self.render_modes = set(CANVAS_RENDER_MODES)

    def _composite(self, triad_data, opts, B, want_f32=True, crop_margin=None):
        uvs = triad_data['uvs'].contiguous()
        default_colors = _unit_range_colors(triad_data['colors'])
        colors = opts.prepare_colors(default_colors, owned=True).contiguous()
        alpha = triad_data['alpha'].contiguous()                     # [B,2,W,W]; channel 0 is alpha_fg
        gen_canvas = triad_data['canvas'].contiguous()
        W = self.patch_width
        out_f32 = torch.empty((B, 4, W, W), dtype=torch.float32, device=self.device) if want_f32 else None
        out_u8, m = None, 0
        if crop_margin is not None:
            m = int(crop_margin)
            out_u8 = torch.empty((B, W - 2 * m, W - 2 * m, 4), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.call('nbe_canvas_composite', _lib.ptr(uvs), _lib.ptr(colors), _lib.ptr(alpha), 2 * W * W, _lib.ptr(gen_canvas),
                      CANVAS_RENDER_MODES[self.render_mode], _lib.ptr(out_f32), _lib.ptr(out_u8), B, W, W, m, _lib.stream())
        return out_f32, out_u8


class InteractiveSession:
    """One-patch-at-a-time rendering for the interactive UI (forger/ui/util.py:176-195 calls ``helper.render_stroke`` per
    websocket message): the whole batch-1 forward -- uint8 patch -> geometry -> encoder -> synthesis -> composite -> uint8
    RGBA -- is captured ONCE into a CUDA graph for the current brush (style, colours, render mode) and replayed per stroke
    patch, so a call costs two small pinned copies and one graph launch instead of ~60 kernel launches issued from Python.

    The graph is re-captured when the brush changes (``set_brush``); stroke geometry and canvas position are the only
    per-call inputs and live in static device buffers."""

    def __init__(self, engine: 'TriadPaintEngine', opts: GanBrushOptions, crop_margin: int = 0):
        self.engine = engine
        self.crop_margin = int(crop_margin)
        dev = engine.device
        W = engine.patch_width
        self._h_patch = torch.empty((W, W), dtype=torch.uint8, pin_memory=True)
        self._h_pos = torch.empty((1, 2), dtype=torch.int64, pin_memory=True)
        self._d_patch = torch.empty((W, W), dtype=torch.uint8, device=dev)
        self._d_pos = torch.zeros((1, 2), dtype=torch.int64, device=dev)
        self._crop0 = torch.zeros((1, 2), dtype=torch.int32, device=dev)
        T = W - 2 * self.crop_margin
        self._h_out = torch.empty((1, T, T, 4), dtype=torch.uint8, pin_memory=True)
        self._stream = torch.cuda.Stream(device=dev)
        self._graph = None
        self.set_brush(opts)

    def _forward(self):
        eng, W = self.engine, self.engine.patch_width
        geom = torch.empty((1, 1, W, W), dtype=torch.float32, device=eng.device)
        # stroke alpha (255 = stroke) -> geometry (0 = stroke): the gather kernel computes 1 - (255 - v) / 255 on v = 255 - alpha
        inv = 255 - self._d_patch
        _lib.call('nbe_gather_geom_patches', _lib.ptr(inv), W, W, _lib.ptr(self._crop0), _lib.ptr(geom), 1, W, _lib.stream())
        self._opts.position = self._d_pos
        tiles, _ = eng.render_tiles(geom, self._opts, crop_margin=self.crop_margin)
        return tiles

    def set_brush(self, opts: GanBrushOptions):
        """(Re)capture the graph for ``opts`` (style z / w+, colours, UVS mapping flag) and the engine's render mode."""
        eng = self.engine
        o = GanBrushOptions()
        o.__dict__.update(opts.__dict__)
        o.to(eng.device)
        for name in ('color0', 'color1', 'canvas_color'):
            c = getattr(o, name, None)
            if c is not None:
                setattr(o, name, c.to(eng.device))
        self._opts = o
        with torch.no_grad(), torch.cuda.device(eng.device):
            if o.enable_uvs_mapping:
                eng.uvs_mapper.get_sfactor(o)                      # fills the per-style cache outside the capture
            self._d_patch.zero_()
            self._stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._stream):
                for _ in range(2):                                  # warm-up: lazy allocations, cudaFuncSetAttribute, workspaces
                    eng.G._noise_cache = None
                    self._forward()
            self._stream.synchronize()
            eng.G._noise_cache = None                               # the noise maps depend on the position: they must be IN the graph
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self._stream):
                self._d_out = self._forward()
            eng.G._noise_cache = None
            self._graph = g
            # the captured kernels hold raw pointers into the batch-1 workspaces of the generator and the encoder: keep them
            # alive (and out of reach of a re-allocation) for as long as the graph exists, whatever other batch sizes the
            # engine serves in between
            self._workspaces = (eng.G._flat_ws.get(1), eng.encoder._ws.get((1, eng.patch_width)), eng.encoder.graph_workspaces(1, eng.patch_width))

    def render_stroke(self, stroke_patch: np.ndarray, position_yx=None) -> np.ndarray:
        """[W,W,C] uint8 stroke patch (last channel: 255 = stroke), optional canvas position (y, x) -> [T,T,4] uint8 RGBA."""
        a = np.ascontiguousarray(stroke_patch[:, :, -1])
        self._h_patch.numpy()[...] = a
        self._h_pos[0, 0], self._h_pos[0, 1] = (0, 0) if position_yx is None else (int(position_yx[0]), int(position_yx[1]))
        with torch.cuda.stream(self._stream):
            self._d_patch.copy_(self._h_patch, non_blocking=True)
            self._d_pos.copy_(self._h_pos, non_blocking=True)
            self._graph.replay()
            self._h_out.copy_(self._d_out, non_blocking=True)
        self._stream.synchronize()
        return self._h_out[0].numpy().copy()
