"""Interactive drawing path of the reference, transport-free.

The reference serves one websocket per user (``forger/ui/run.py:58-142``); every binary message carries one 128^2 stroke
patch, is decoded by ``forger/ui/util.py:50-104``, rendered by ``PaintingHelper.render_stroke``
(``forger/ui/brush.py:244-398``) and answered with one binary message (``util.py:26-47,161-195``).  This module holds
everything of that path except the Tornado / Flask plumbing (not installable here, and not on the hot path):

* the wire codec -- same function names, argument meaning and byte layout as ``forger/ui/util.py:21-104``, plus the client's
  side of it (``encode_render_request`` / ``decode_render_response`` restate ``forger/ui/js/main_controller.js:532-677``)
  so that tests and non-browser clients can speak the protocol;
* ``PaintingHelper`` -- per-session brush, render mode, geometry canvas size and feature canvas (``brush.py:95-398``), on
  the B200 engine;
* ``DrawingSession`` -- ``DrawingWebSocketHandler`` (``util.py:107-245``) as a plain object: ``open()`` and
  ``on_message(msg)`` return the messages the handler would write to its socket;
* ``StrokeBatcher`` -- what the reference does not have: render requests of CONCURRENT sessions are collected and rendered
  as ONE batched forward (every patch with its own style, colours, position and feature-canvas window), because at batch
  1 the B200 is launch-latency-bound (0.36-0.9 ms per patch) while a batch of 32 costs barely more than one patch.

There is no CPU path: everything renders through ``TriadPaintEngine`` (libnbe_b200).
"""
from __future__ import annotations

import copy
import json
import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .engine import GanBrushOptions, TriadPaintEngine
from .stylizer import RasterFeatureCanvas, dirty_area_alpha, _flat_blend_ok

RESPONSE_RENDER, RESPONSE_DEBUG_IMAGE, RESPONSE_BRUSH_SAMPLE = 0, 1, 2       # util.py:161-171


# ------------------------------------------------------------------------------------------------ wire codec
def int32_to_binary(single_int) -> bytes:
    """util.py:21-22."""
    return np.array([single_int], dtype=np.int32).tobytes()


def image_patch_to_binary(img: np.ndarray, x, y) -> bytes:
    """``int32 width, height, x, y`` + ``height * width * 4`` uint8 RGBA (util.py:26-47)."""
    if img.dtype != np.uint8:
        raise RuntimeError('Image must be uint8 in range 0...255')
    height, width, nchannels = img.shape[0], img.shape[1], img.shape[2]
    assert nchannels < height, f'Wrong shape {img.shape}'
    return np.array([width, height, x, y], dtype=np.int32).tobytes() + img.tobytes()


def binary_to_image_patches(bytes_msg: bytes, offset: int = 0):
    """``int32 width, height, x, y, crop_margin`` + RGBA stroke canvas -> (meta, stroke image [h,w,4] uint8, None)
    (util.py:50-76; the reference never decodes the second, canvas image either)."""
    metadata = np.frombuffer(bytes_msg, dtype=np.int32, count=5, offset=offset)
    meta = {'width': metadata[0], 'height': metadata[1], 'x': metadata[2], 'y': metadata[3], 'crop_margin': metadata[4]}
    img_data = np.frombuffer(bytes_msg, dtype=np.uint8, offset=offset + 5 * 4)
    imgsize = int(meta['height']) * int(meta['width']) * 4
    img_stroke = img_data[0:imgsize].reshape((meta['height'], meta['width'], 4))
    return meta, img_stroke, None


def decode_render_request_metadata(bytes_msg: bytes, offset: int = 0):
    """``uint8 debug, ncolors, extra`` + ``ncolors x (uint8 idx, R, G, B)`` -> (meta, next read offset) (util.py:79-104)."""
    metadata = np.frombuffer(bytes_msg, dtype=np.uint8, count=3, offset=offset)
    read_start = offset + 3
    meta = {'debug': metadata[0] != 0, 'colors': [], 'extra_data': metadata[2]}
    for _ in range(int(metadata[1])):
        meta['colors'].append(np.frombuffer(bytes_msg, dtype=np.uint8, count=4, offset=read_start))
        read_start += 4
    return meta, read_start


def encode_render_request(stroke_rgba: np.ndarray, x: int, y: int, crop_margin: int = 0,
                          colors: Sequence[Tuple[int, int, int, int]] = (), debug: bool = False, extra_data: int = 0) -> bytes:
    """The client's request (``main_controller.js:532-657``): stroke_rgba [h,w,4] uint8 with alpha = stroke geometry;
    ``colors`` = (colour index 0/1/2, R, G, B) overrides."""
    if stroke_rgba.dtype != np.uint8 or stroke_rgba.ndim != 3 or stroke_rgba.shape[2] != 4:
        raise RuntimeError('stroke image must be [h, w, 4] uint8')
    head = np.array([int(bool(debug)), len(colors), int(extra_data)], dtype=np.uint8).tobytes()
    for c in colors:
        head += np.array(c, dtype=np.uint8).tobytes()
    h, w = stroke_rgba.shape[:2]
    return head + np.array([w, h, x, y, crop_margin], dtype=np.int32).tobytes() + np.ascontiguousarray(stroke_rgba).tobytes()


def decode_render_response(bytes_msg: bytes):
    """``int32 type`` + image patch (``main_controller.js:664-677``) -> (type, {'width','height','x','y'}, img [h,w,4])."""
    head = np.frombuffer(bytes_msg, dtype=np.int32, count=5)
    meta = {'width': int(head[1]), 'height': int(head[2]), 'x': int(head[3]), 'y': int(head[4])}
    img = np.frombuffer(bytes_msg, dtype=np.uint8, offset=20, count=meta['width'] * meta['height'] * 4)
    return int(head[0]), meta, img.reshape(meta['height'], meta['width'], 4)


# ------------------------------------------------------------------------------------------------ feature canvases
class FeaturePool:
    """Device memory behind the feature canvases of ALL sessions at one blending level on the flat bf16 path: one NHWC
    bf16 tensor [rows, cols, C] + uint8 mask; a session's canvas is a band of rows.  Because every window of every session
    then lives in the same tensor, one ``generator.WindowBlend`` (one ``nbe_blend_window_nhwc_bf16`` launch) serves a batch
    that mixes sessions.  Bands are zeroed when handed out; ``release`` returns them."""

    def __init__(self, engine: TriadPaintEngine, level: int, rows: int, cols: int):
        self.engine, self.level = engine, int(level)
        self.down = 2 ** (level - 1)
        self.res = engine.patch_width // self.down
        C = engine.G.cfg.channels(self.res)
        self.fcanvas = torch.zeros((rows, cols, C), dtype=torch.bfloat16, device=engine.device)
        self.fmask = torch.zeros((rows, cols), dtype=torch.uint8, device=engine.device)
        self._free: List[Tuple[int, int]] = [(0, rows)]                 # (first row, number of rows), sorted
        self._alpha: Dict[int, torch.Tensor] = {}

    def base_alpha(self, cm: int) -> torch.Tensor:
        a = self._alpha.get(cm)
        if a is None:
            margin = 16 // self.down                                    # PaintingHelper.feature_blending_margin = 16
            a = self._alpha[cm] = dirty_area_alpha(self.res, margin, cm, self.engine.device).to(torch.float32).contiguous()
        return a

    def acquire(self, fh: int, fw: int) -> int:
        """First row of a zeroed band that holds an fh x fw feature canvas plus the ``res`` rows / columns a window at its
        bottom / right edge reaches into."""
        need = fh + self.res
        if fw + self.res > self.fcanvas.shape[1]:
            raise RuntimeError(f'FeaturePool: canvas of {fw} feature columns does not fit the pool ({self.fcanvas.shape[1] - self.res})')
        for i, (r0, n) in enumerate(self._free):
            if n >= need:
                self._free[i:i + 1] = [(r0 + need, n - need)] if n > need else []
                self.fcanvas[r0:r0 + need].zero_()
                self.fmask[r0:r0 + need].zero_()
                return r0
        raise RuntimeError(f'FeaturePool: no band of {need} rows left (level {self.level})')

    def release(self, r0: int, fh: int):
        self._free.append((r0, fh + self.res))
        self._free.sort()
        merged: List[Tuple[int, int]] = []
        for a, n in self._free:
            if merged and merged[-1][0] + merged[-1][1] == a:
                merged[-1] = (merged[-1][0], merged[-1][1] + n)
            else:
                merged.append((a, n))
        self._free = merged


def _engine_pool(engine: TriadPaintEngine, level: int, fh: int, fw: int) -> FeaturePool:
    """The engine-wide pool of a blending level, created on first use: at least 2048 canvas pixels wide (or the first
    requested width if larger), and as many rows as ``engine.feature_pool_bytes`` (default 1 GiB) buys."""
    pools = engine.__dict__.setdefault('_feature_pools', {})
    pool = pools.get(level)
    if pool is None:
        down = 2 ** (level - 1)
        res = engine.patch_width // down
        C = engine.G.cfg.channels(res)
        cols = max(fw, 2048 // down) + res
        budget = int(getattr(engine, 'feature_pool_bytes', 1 << 30))
        rows = max(fh + res, budget // (cols * C * 2))
        pool = pools[level] = FeaturePool(engine, level, rows, cols)
    return pool


# ------------------------------------------------------------------------------------------------ PaintingHelper
class PaintingHelper:
    """``forger/ui/brush.py:95-398``: brush, canvas and feature-canvas state of one session.  ``render_stroke`` keeps the
    reference's signature and results (uint8 [W-2m, W-2m, 4] patch, debug image, ``{'x','y'}`` of its top-left corner)."""

    def __init__(self, paint_engine: TriadPaintEngine, style_seed=None, debug_dir=None):
        self.engine = paint_engine
        self.seed_rng = np.random.default_rng(seed=style_seed)
        self.brush_options = GanBrushOptions()
        self.brush_options.set_style(*self.random_brush_style())
        self.debug_dir = debug_dir
        self.render_id = 0
        self.canvas_shape: Optional[Tuple[int, int]] = None              # the reference keeps a float geometry canvas it never reads
        self.feature_blending_level = 0
        self.feature_blending_margin = 16
        self._raster: Optional[RasterFeatureCanvas] = None               # generic (FP32 / non-stock) feature canvas
        self._pool: Optional[FeaturePool] = None                         # flat bf16 path: band of the engine's pool
        self._band = None

    # ---- canvas ----
    def make_new_canvas(self, rows, cols, feature_blending=None):
        self.canvas_shape = (int(rows), int(cols))
        self.set_feature_blending(self.feature_blending_level if feature_blending is None else feature_blending)

    def set_feature_blending(self, feature_blending_level=0):
        self.close()
        self.feature_blending_level = int(feature_blending_level)
        if self.feature_blending_level <= 0:
            return
        if self.canvas_shape is None:
            raise RuntimeError('make_new_canvas must be called before feature blending is enabled')
        down = 2 ** (self.feature_blending_level - 1)
        fh, fw = int(math.ceil(self.canvas_shape[0] / down)), int(math.ceil(self.canvas_shape[1] / down))
        if _flat_blend_ok(self.engine):
            self._pool = _engine_pool(self.engine, self.feature_blending_level, fh, fw)
            self._band = (self._pool.acquire(fh, fw), fh, fw)
        else:
            self._raster = RasterFeatureCanvas(self.engine, self.feature_blending_level, fh, fw, self.feature_blending_margin)

    def close(self):
        """Give the feature canvas back (a session that ends, or a new canvas)."""
        if self._pool is not None and self._band is not None:
            self._pool.release(self._band[0], self._band[1])
        self._pool, self._band, self._raster = None, None, None

    @property
    def down_factor(self) -> int:
        return 2 ** (self.feature_blending_level - 1) if self.feature_blending_level > 0 else 1

    # ---- brush ----
    def set_new_brush(self, seed=None):
        style_z, seed = self.random_brush_style(seed)
        self.brush_options.set_style(style_z, seed)
        return seed

    def set_render_mode(self, mode=None):
        self.engine.set_render_mode(mode)

    def generate_style_seed(self):
        return self.seed_rng.integers(low=0, high=10000, size=1)[0]

    def random_brush_style(self, seed=None):
        if seed is None:
            seed = self.generate_style_seed()
        return self.engine.random_style(seed), seed

    def default_brush_options(self):
        return copy.copy(self.brush_options)                             # shallow, as in the reference

    # ---- rendering ----
    def snap(self, meta) -> Tuple[Optional[int], Optional[int], int]:
        """(y, x) snapped down to the feature grid and the crop margin of a request (brush.py:250-268)."""
        if meta is None:
            return None, None, 0
        x, y = int(meta.get('x')), int(meta.get('y'))
        d = self.down_factor
        x, y = (x // d) * d, (y // d) * d
        return y, x, int(meta.get('crop_margin')) if 'crop_margin' in meta else 0

    def _check_window(self, y, x):
        if self.feature_blending_level > 0:
            assert y is not None, 'feature blending needs the x, y of the patch'
            H, W = self.canvas_shape
            if not (0 <= y < H and 0 <= x < W):
                raise RuntimeError(f'patch at ({y}, {x}) starts outside the {H}x{W} canvas')

    def render_stroke(self, stroke_patch, canvas_patch, opts, meta=None):
        H, W, _ = stroke_patch.shape
        if W != self.engine.patch_width or H != self.engine.patch_width:
            raise RuntimeError('Not implemented')                        # brush.py:277-278
        y, x, crop_margin = self.snap(meta)
        self._check_window(y, x)
        geom = self.engine.prepare_geom_input(stroke_patch)
        with torch.no_grad(), torch.cuda.device(self.engine.device):
            if self._raster is not None:
                tiles = self._raster.render(geom, opts, y, x, crop_margin)
            elif self._band is not None:
                d = self.down_factor
                fyx = torch.tensor([[self._band[0] + y // d, x // d]], dtype=torch.int32, device=self.engine.device)
                from .generator import WindowBlend
                cm = crop_margin // d
                wb = WindowBlend(self._pool.res, self._pool.fcanvas, self._pool.fmask, fyx, self._pool.base_alpha(cm), cm)
                tiles, _ = self.engine.render_tiles(geom, opts, crop_margin=crop_margin, window_blend=wb)
            else:
                tiles, _ = self.engine.render_tiles(geom, opts, crop_margin=crop_margin)
        img = np.ascontiguousarray(tiles[0].cpu().numpy())
        out_meta = {'x': (x if x is not None else 0) + crop_margin, 'y': (y if y is not None else 0) + crop_margin}
        return img, None, out_meta


# ------------------------------------------------------------------------------------------------ batching across sessions
class _BatchOptions(GanBrushOptions):
    """Brush options of a batch whose patches come from different sessions: per-patch colour overrides and UVS factors."""

    def __init__(self):
        super().__init__()
        self.color_mask = None          # [B, 3] bool: column idx of patch b is overridden
        self.color_vals = None          # [B, 3 (rgb), 3 (idx)]
        self.sfactor = None             # [B] float32 (UVS mapping on) or None

    def prepare_colors(self, default_colors, owned=False):
        if self.color_mask is None:
            return default_colors if owned else default_colors.clone()
        return torch.where(self.color_mask[:, None, :], self.color_vals, default_colors)


class _Ticket:
    __slots__ = ('helper', 'geom_u8', 'opts', 'y', 'x', 'crop_margin', 'result')

    def __init__(self, helper, geom_u8, opts, y, x, crop_margin):
        self.helper, self.geom_u8, self.opts, self.y, self.x, self.crop_margin = helper, geom_u8, opts, y, x, crop_margin
        self.result = None


class StrokeBatcher:
    """Collects ``render_stroke`` requests of concurrent sessions (``submit``) and renders them with as few generator
    forwards as the requests allow (``flush``): requests are grouped by what must be uniform inside one forward (crop
    margin, with / without canvas positions, UVS mapping on / off, blending level), every patch keeps its own style,
    colours, position and feature-canvas window.  Two requests of the SAME session with feature blending are
    raster-dependent (the second reads what the first saved) and go into successive waves.  Results equal the
    one-at-a-time ``PaintingHelper.render_stroke`` byte for byte: every kernel on the path is batch-invariant.

    Batched: plain z / w+ styles on the flat bf16 path.  Rendered one by one through their own helper: brushes with
    per-brush noise buffers (projected w+ libraries), FP32 engines with feature blending."""

    def __init__(self, engine: TriadPaintEngine, max_batch: int = 64):
        self.engine, self.max_batch = engine, int(max_batch)
        self._pending: List[_Ticket] = []
        self.forwards = 0               # generator forwards issued so far (what batching saves)

    def submit(self, helper: PaintingHelper, stroke_patch: np.ndarray, opts: GanBrushOptions, meta=None) -> _Ticket:
        H, W, _ = stroke_patch.shape
        if W != self.engine.patch_width or H != self.engine.patch_width:
            raise RuntimeError('Not implemented')
        y, x, crop_margin = helper.snap(meta)
        helper._check_window(y, x)
        t = _Ticket(helper, np.ascontiguousarray(stroke_patch[:, :, -1]), opts, y, x, crop_margin)
        self._pending.append(t)
        return t

    def _batchable(self, t: _Ticket) -> bool:
        o = t.opts
        if o.style_ws is not None and (o.custom_args or {}).get('noise_buffers') is not None:
            return False
        if t.helper.feature_blending_level > 0 and t.helper._band is None:
            return False
        return self.engine.G.flat_supported and self.engine.encoder.mode == 'bf16'

    def flush(self) -> List[_Ticket]:
        """Render everything submitted since the last flush; returns the tickets (``ticket.result`` =
        (img, None, out_meta)) in submission order."""
        pending, self._pending = self._pending, []
        groups: Dict[tuple, List[List[_Ticket]]] = {}
        for t in pending:
            if not self._batchable(t):
                self.forwards += 1
                meta = None if t.y is None else {'x': t.x, 'y': t.y, 'crop_margin': t.crop_margin}
                rgba = np.zeros(t.geom_u8.shape + (4,), dtype=np.uint8)
                rgba[..., 3] = t.geom_u8
                t.result = t.helper.render_stroke(rgba, None, t.opts, meta)
                continue
            key = (t.crop_margin, t.opts.position is not None, bool(t.opts.enable_uvs_mapping), t.helper.feature_blending_level)
            waves = groups.setdefault(key, [])
            # the k-th request of a blending session goes to wave k; everything else to wave 0
            k = 0
            if t.helper.feature_blending_level > 0:
                k = sum(1 for w in waves for u in w if u.helper is t.helper)
            while len(waves) <= k:
                waves.append([])
            waves[k].append(t)
        for key, waves in groups.items():
            for wave in waves:
                for i in range(0, len(wave), self.max_batch):
                    self._render(key, wave[i:i + self.max_batch])
        return pending

    def _render(self, key, tickets: List[_Ticket]):
        crop_margin, has_pos, uvs_mapping, level = key
        eng, dev, W = self.engine, self.engine.device, self.engine.patch_width
        n = len(tickets)
        self.forwards += 1
        with torch.no_grad(), torch.cuda.device(dev):
            alpha = torch.from_numpy(np.stack([t.geom_u8 for t in tickets])).to(dev)
            geom = (1 - alpha.to(torch.float32) / 255.0)[:, None]                                # prepare_geom_input, batched
            o = _BatchOptions()
            if any(t.opts.style_ws is not None for t in tickets):
                ws = []
                for t in tickets:
                    if t.opts.style_ws is not None:
                        ws.append(eng.G.expand_ws(t.opts.style_ws.to(dev, torch.float32))[:1])
                    else:
                        ws.append(eng.G.mapping(t.opts.style_z.to(dev)[:1], eng.style_c).to(torch.float32))
                o.set_style_w(torch.cat(ws).contiguous())
            else:
                o.set_style(torch.cat([t.opts.style_z.to(dev)[:1] for t in tickets]))
            if has_pos:
                o.position = torch.cat([t.opts.position.to(dev)[:1] for t in tickets])
            mask = torch.zeros((n, 3), dtype=torch.bool)
            vals = torch.zeros((n, 3, 3), dtype=torch.float32)
            for b, t in enumerate(tickets):
                for idx, col in enumerate((t.opts.color0, t.opts.color1, t.opts.canvas_color)):
                    if col is not None:
                        mask[b, idx] = True
                        vals[b, :, idx] = col.reshape(-1)[:3].cpu()
            if bool(mask.any()):
                o.color_mask, o.color_vals = mask.to(dev), vals.to(dev)
            if uvs_mapping:
                o.enable_uvs_mapping = True
                sf = eng.uvs_mapper.get_sfactors([t.opts for t in tickets])
                o.sfactor = torch.stack([s.reshape(()) for s in sf]).to(dev, torch.float32)
            kw = {}
            if level > 0:
                from .generator import WindowBlend
                pool = tickets[0].helper._pool
                d = pool.down
                fyx = torch.tensor([[t.helper._band[0] + t.y // d, t.x // d] for t in tickets], dtype=torch.int32, device=dev)
                cm = crop_margin // d
                kw['window_blend'] = WindowBlend(pool.res, pool.fcanvas, pool.fmask, fyx, pool.base_alpha(cm), cm)
            tiles, _ = eng.render_tiles(geom, o, crop_margin=crop_margin, **kw)
            tiles = tiles.cpu().numpy()
        for b, t in enumerate(tickets):
            out_meta = {'x': (t.x if t.x is not None else 0) + crop_margin, 'y': (t.y if t.y is not None else 0) + crop_margin}
            t.result = (np.ascontiguousarray(tiles[b]), None, out_meta)


# ------------------------------------------------------------------------------------------------ session (websocket handler)
class DrawingSession:
    """``DrawingWebSocketHandler`` (``forger/ui/util.py:107-245``) without the socket: ``open()`` / ``on_message(msg)``
    return the list of messages the handler would have written, each ``(payload, binary)`` with JSON payloads as dicts.
    With a ``batcher`` the binary render requests are only queued (``on_message`` returns ``[]`` for them) and answered by
    ``DrawingSession.flush_all(batcher)`` -- one batched forward for all sessions that sent a stroke in the meantime.
    ``opts.debug`` is accepted and ignored: the didactic debug image (``brush.py:807-868``) is visualisation, out of scope."""

    def __init__(self, paint_engine: TriadPaintEngine, style_seed=None, debug_dir=None, saved_zs_filename=None, libraries=None,
                 batcher: Optional[StrokeBatcher] = None):
        self.helper = PaintingHelper(paint_engine, style_seed=style_seed, debug_dir=debug_dir)
        self.zs_file = saved_zs_filename
        self.libraries = libraries if libraries is not None else {}
        self.use_positions = False
        self.uvs_mapping = False
        self.batcher = batcher
        self._queued: List[Tuple[_Ticket, int]] = []

    def open(self):
        return [({'type': 'modelinfo', 'data': {'patch_width': self.helper.engine.patch_width}}, False), self._brush_info()]

    def close(self):
        self.helper.close()

    def _brush_info(self):
        o = self.helper.brush_options
        return ({'type': 'brushinfo', 'data': {'style_id': '%s' % str(o.style_id), 'library_id': '%s' % o.library_id,
                                               'colors': '%s' % self.helper.engine.uvs_mapper.get_colors(o)}}, False)

    def save_current_brush(self):
        o = self.helper.brush_options
        if self.zs_file is None or o.style_id is None:
            return
        with open(self.zs_file, 'a') as f:
            f.write(('%d ' % o.style_id) + ' '.join('%f' % v for v in o.style_z[0, ...].tolist()) + '\n')

    def on_message(self, message):
        """Never raises: like the reference handler, a message that cannot be decoded or rendered is dropped."""
        try:
            if isinstance(message, (bytes, bytearray, memoryview)):
                return self._handle_binary_request(bytes(message))
            return self._handle_json_request(message)
        except Exception as e:                                            # noqa: BLE001  (util.py:153-159)
            self.last_error = e
            return []

    @staticmethod
    def _encode_type_render(extra_data):
        return int32_to_binary(0 if extra_data == 0 else extra_data)

    def _request_options(self, meta) -> GanBrushOptions:
        o = self.helper.default_brush_options()
        for colorinfo in meta['colors']:
            o.set_color(int(colorinfo[0]), colorinfo[1:])
        o.debug = bool(meta['debug'])
        if self.use_positions:
            o.set_position(int(meta['x']), int(meta['y']))
        else:
            o.position = None
        o.enable_uvs_mapping = self.uvs_mapping
        return o

    def _handle_binary_request(self, raw_message: bytes):
        meta, read_offset = decode_render_request_metadata(raw_message)
        patch_meta, img_stroke, img_canvas = binary_to_image_patches(raw_message, read_offset)
        meta.update(patch_meta)
        o = self._request_options(meta)
        if self.batcher is not None:
            self._queued.append((self.batcher.submit(self.helper, img_stroke, o, meta), int(meta['extra_data'])))
            return []
        res_img, debug_img, meta_out = self.helper.render_stroke(img_stroke, img_canvas, o, meta)
        return [(self._encode_type_render(int(meta['extra_data'])) + image_patch_to_binary(res_img, meta_out['x'], meta_out['y']), True)]

    def collect(self):
        """Responses of the queued render requests after ``batcher.flush()``."""
        out = []
        for t, extra in self._queued:
            if t.result is not None:
                img, _, m = t.result
                out.append((self._encode_type_render(extra) + image_patch_to_binary(img, m['x'], m['y']), True))
        self._queued = []
        return out

    @staticmethod
    def flush_all(batcher: StrokeBatcher, sessions: Sequence['DrawingSession']):
        """One batched render of everything the sessions queued -> {session: [messages]}."""
        batcher.flush()
        return {s: s.collect() for s in sessions}

    def _handle_json_request(self, raw_message):
        msg = json.loads(raw_message) if isinstance(raw_message, str) else raw_message
        kind = msg.get('type')
        if kind == 'set_brush':
            if msg.get('style_id') and msg.get('library_id'):
                library_id, style_id = msg.get('library_id'), msg.get('style_id')
                if library_id in self.libraries and style_id in self.libraries[library_id].get_style_ids():
                    self.libraries[library_id].set_style(style_id, self.helper.brush_options)
                    self.helper.brush_options.library_id = library_id
            else:
                self.helper.set_new_brush(msg.get('seed'))
            return [self._brush_info()]
        if kind == 'save_brush':
            self.save_current_brush()
        elif kind == 'set_option':
            if msg.get('option') == 'positions':
                self.use_positions = msg.get('value')
            elif msg.get('option') == 'uvs_mapping':
                self.uvs_mapping = msg.get('value')
        elif kind == 'set_render_mode':
            self.helper.set_render_mode(msg.get('mode'))
        elif kind == 'new_canvas':
            self.helper.make_new_canvas(int(msg.get('rows')), int(msg.get('cols')), feature_blending=int(msg.get('feature_blending')))
        return []
