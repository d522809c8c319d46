"""Patch scheduler: the batched, multi-GPU replacement for the per-patch loop of
forger/viz/paint_image_main.py:145-186 (+ forger/viz/style_transfer.py:15-48).

Semantics kept bit-exact (integer work): crop list, padded canvas size, ``meta`` offsets ``(y+m, x+m)``,
last-writer-wins tile placement in raster order, final crop back to the input size, ``on_white`` composite.

Execution on B200:
* the padded guidance image lives on the device once (uint8); each batch of crops is gathered straight into
  the encoder's float input (``nbe_gather_geom_patches`` = crop + ``255 - g`` + ``prepare_geom_input``);
* patches of a batch go through encoder -> generator -> composite as ONE batched launch sequence
  (``TriadPaintEngine.render_tiles``), per-patch ``positions`` = crop (y, x) exactly as the loop sets them;
* tiles are written into a device canvas under an ownership map (``nbe_tile_owner_map`` / ``nbe_place_tiles``),
  so overlapping tiles of one batch resolve to "highest raster index wins" without ordering the launches;
* across GPUs, contiguous bands of crop rows go to ranks (patches are independent when
  ``feature_blending_level == 0``); the only inter-GPU traffic is one gather of finished uint8 tiles to rank 0.

``feature_blending_level > 0`` makes patch n read features written by raster-earlier neighbours
(forger/ui/brush.py:190-242); it is executed in raster order with the reference's exact mask arithmetic
(single GPU), see ``_stylize_blended``.
"""
from __future__ import annotations

import math
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .engine import GanBrushOptions, TriadPaintEngine


# ------------------------------------------------------------------------------------------------ integer host logic
def pad_geo(geo: np.ndarray, crop_margin: int) -> np.ndarray:
    """paint_image_main.py:59-62."""
    out = np.full((geo.shape[0] + crop_margin, geo.shape[1] + crop_margin, geo.shape[2]), 255, dtype=np.uint8)
    out[crop_margin:, crop_margin:, :] = geo
    return out


def generate_stitching_crops(stroke_image: np.ndarray, patch_width: int, mode: str = 'all', overlap_margin: int = 15):
    """style_transfer.py:15-48 -> (list of (y, x, w, w), padded image)."""
    rwidth = patch_width - overlap_margin * 2
    img_height, img_width, nchannels = stroke_image.shape
    assert nchannels in [1, 2, 3, 4], f'Wrong shape {stroke_image.shape}'
    nrows = img_height // rwidth + 1
    ncols = img_width // rwidth + 1
    padded = np.full((nrows * rwidth + patch_width, ncols * rwidth + patch_width, nchannels), 255, dtype=np.uint8)
    padded[0:img_height, 0:img_width, ...] = stroke_image
    if mode == 'all':
        ys, xs = np.meshgrid(np.arange(nrows) * rwidth, np.arange(ncols) * rwidth, indexing='ij')
        crops = [(int(y), int(x), patch_width, patch_width) for y, x in zip(ys.ravel(), xs.ravel())]
    else:
        crops = []
        for r in range(nrows):
            for c in range(ncols):
                y, x = r * rwidth, c * rwidth
                if np.sum(padded[y:y + patch_width, x:x + patch_width, ...] < 0.001) > 10:
                    crops.append((y, x, patch_width, patch_width))
    return crops, padded


def shard_crops(crops: Sequence[Tuple[int, int, int, int]], world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous bands of crop *rows* per rank: -> [start, end) indices into ``crops``.
    Rows are split as evenly as possible, earlier ranks take the extra rows (47 rows / 8 -> 6,6,6,6,6,6,6,5)."""
    row_ys = sorted({c[0] for c in crops})
    nrows = len(row_ys)
    base, extra = divmod(nrows, world_size)
    r0 = rank * base + min(rank, extra)
    r1 = r0 + base + (1 if rank < extra else 0)
    if r0 >= nrows:
        return len(crops), len(crops)
    ys = [c[0] for c in crops]
    start = ys.index(row_ys[r0])
    end = len(crops) if r1 >= nrows else ys.index(row_ys[r1])
    return start, end


def shard_bounds(crops_yx: np.ndarray, world_size: int) -> List[Tuple[int, int]]:
    """``shard_crops`` for every rank at once from the [n, 2] (y, x) array -- one vectorised pass instead of ``world_size``
    Python scans of the crop list (those scans were ~3 ms of the 8-GPU canvas latency)."""
    n = len(crops_yx)
    if n == 0:
        return [(0, 0)] * world_size
    row_ys, first = np.unique(crops_yx[:, 0], return_index=True)       # raster order: rows are contiguous, ys ascending
    nrows = len(row_ys)
    base, extra = divmod(nrows, world_size)
    out = []
    for rank in range(world_size):
        r0 = rank * base + min(rank, extra)
        r1 = r0 + base + (1 if rank < extra else 0)
        if r0 >= nrows:
            out.append((n, n))
        else:
            out.append((int(first[r0]), n if r1 >= nrows else int(first[r1])))
    return out


def composite_on_white(result: np.ndarray) -> np.ndarray:
    """paint_image_main.py:179-183."""
    alpha = result[..., 3:].astype(np.float32) / 255
    out = result[..., :3].astype(np.float32) * alpha + 255 * (1 - alpha)
    return out.clip(0, 255).astype(np.uint8)


# ------------------------------------------------------------------------------------------------ device-side scheduler
def crop_grid(img_height: int, img_width: int, patch_width: int, overlap_margin: int):
    """Arithmetic of style_transfer.py:23-30: -> (nrows, ncols, rwidth, padded_h, padded_w)."""
    rwidth = patch_width - overlap_margin * 2
    nrows = img_height // rwidth + 1
    ncols = img_width // rwidth + 1
    return nrows, ncols, rwidth, nrows * rwidth + patch_width, ncols * rwidth + patch_width


class CanvasJob:
    """One stylization of one guidance image on one rank.  ``guidance`` may be a host ``np.ndarray`` or a CUDA uint8
    tensor [H,W,C]; the padded canvas (pad_geo + generate_stitching_crops padding, value 255) is built on the device
    and the crop list comes from the grid arithmetic, so nothing but the (optional) initial upload touches the host."""

    def __init__(self, engine: TriadPaintEngine, guidance, crop_margin: int = 10, stitching_mode: str = 'all',
                 crop_rows: Optional[Tuple[int, int]] = None):
        """``crop_rows = (r0, r1)`` (dense 'all' grid only) restricts the job to the crop rows [r0, r1) of the grid: only the
        guidance rows those crops read are uploaded / kept on the device, the crop list holds those rows only (global (y, x)
        coordinates, as the generator's noise shift needs them) and ``band`` gives the canvas rows the job's tiles own."""
        self.engine = engine
        dev = engine.device
        self.crop_margin = m = int(crop_margin)
        self.patch = engine.patch_width
        if crop_rows is not None:
            self._init_row_window(guidance, stitching_mode, crop_rows)
            return
        self.geom_row0 = 0
        self.band = None
        extra_channels = []                                              # channels besides the last one (sparse modes count them too)
        if isinstance(guidance, np.ndarray):
            assert guidance.ndim == 3 and guidance.dtype == np.uint8
            if stitching_mode != 'all':
                extra_channels = [torch.from_numpy(np.ascontiguousarray(guidance[:, :, c])).to(dev) for c in range(guidance.shape[2] - 1)]
            guidance = torch.from_numpy(np.ascontiguousarray(guidance[:, :, -1])).to(dev)
        else:
            assert guidance.ndim == 3 and guidance.dtype == torch.uint8
            if stitching_mode != 'all':
                extra_channels = [guidance[:, :, c].to(dev) for c in range(guidance.shape[2] - 1)]
            guidance = guidance[:, :, -1].to(dev)
        H0, W0 = int(guidance.shape[0]), int(guidance.shape[1])
        self.orig_shape = (H0, W0)
        nrows, ncols, rwidth, ph, pw = crop_grid(H0 + m, W0 + m, self.patch, m * 2)
        self.canvas_h, self.canvas_w = ph, pw
        self.d_geom = torch.full((ph, pw), 255, dtype=torch.uint8, device=dev)
        self.d_geom[m:m + H0, m:m + W0] = guidance                       # pad_geo offset + bottom/right padding in one go
        ys, xs = np.meshgrid(np.arange(nrows) * rwidth, np.arange(ncols) * rwidth, indexing='ij')
        yx = np.stack([ys.ravel(), xs.ravel()], axis=1).astype(np.int32)
        d_yx = None
        if stitching_mode != 'all':
            # keep crops with more than 10 stroke (zero) pixels (style_transfer.py:45): window sums on the device
            # (the reference sums over every channel of the stroke image; its own caller always passes one channel)
            counts = torch.empty((nrows * ncols,), dtype=torch.int32, device=dev)
            with torch.cuda.device(dev):
                _lib.call('nbe_count_stroke_pixels', _lib.ptr(self.d_geom), ph, pw, self.patch, rwidth, nrows, ncols, _lib.ptr(counts),
                          _lib.stream())
                for ch in extra_channels:
                    padded = torch.full((ph, pw), 255, dtype=torch.uint8, device=dev)
                    padded[m:m + H0, m:m + W0] = ch
                    more = torch.empty_like(counts)
                    _lib.call('nbe_count_stroke_pixels', _lib.ptr(padded), ph, pw, self.patch, rwidth, nrows, ncols, _lib.ptr(more),
                              _lib.stream())
                    counts += more
            keep = (counts > 10).cpu().numpy()
            yx = yx[keep]
        else:
            # the full grid is pure arithmetic: build it on the device too instead of uploading it (no synchronous copy)
            gy = torch.arange(nrows, dtype=torch.int32, device=dev) * rwidth
            gx = torch.arange(ncols, dtype=torch.int32, device=dev) * rwidth
            d_yx = torch.stack([gy[:, None].expand(nrows, ncols), gx[None, :].expand(nrows, ncols)], dim=2).reshape(-1, 2).contiguous()
        self.crops_yx = np.ascontiguousarray(yx).reshape(-1, 2)
        self._crops = None
        self.tile = self.patch - 2 * m
        self.d_crops = d_yx if d_yx is not None else torch.from_numpy(self.crops_yx).to(dev)
        self.tiles_yx = self.crops_yx + m                                # meta: (y + m, x + m)  (brush.py:365-373)
        self.d_tiles_yx = self.d_crops + m
        self.d_crops_local = self.d_crops
        self.d_tiles_yx_band = None

    def _init_row_window(self, guidance, stitching_mode, crop_rows):
        if stitching_mode != 'all':
            raise RuntimeError("CanvasJob: a crop-row window needs the dense grid (stitching_mode='all')")
        dev, m = self.engine.device, self.crop_margin
        assert guidance.ndim == 3
        H0, W0 = int(guidance.shape[0]), int(guidance.shape[1])
        self.orig_shape = (H0, W0)
        nrows, ncols, rwidth, ph, pw = crop_grid(H0 + m, W0 + m, self.patch, m * 2)
        r0, r1 = int(crop_rows[0]), int(crop_rows[1])
        assert 0 <= r0 <= r1 <= nrows
        self.canvas_h, self.canvas_w = ph, pw
        self.tile = self.patch - 2 * m
        # padded-canvas rows [g0, g1) are what the crops of rows [r0, r1) read; padded row Y is guidance row Y - m
        g0 = r0 * rwidth
        g1 = max(g0, (r1 - 1) * rwidth + self.patch) if r1 > r0 else g0
        self.geom_row0 = g0
        self.d_geom = torch.full((max(g1 - g0, 1), pw), 255, dtype=torch.uint8, device=dev)
        s0, s1 = max(0, g0 - m), min(H0, g1 - m)
        if s1 > s0:
            if isinstance(guidance, np.ndarray):
                assert guidance.dtype == np.uint8
                # pinned staging: the slice is a few MB, the copy is asynchronous and at full PCIe rate
                host = torch.empty((s1 - s0, W0), dtype=torch.uint8, pin_memory=True)
                host.numpy()[...] = guidance[s0:s1, :, -1]
                src = host.to(dev, non_blocking=True)
            else:
                assert guidance.dtype == torch.uint8
                src = guidance[s0:s1, :, -1].to(dev)
            self.d_geom[s0 + m - g0:s1 + m - g0, m:m + W0] = src
        # the crop grid of a row window is pure arithmetic: built once per (shape, margin, rows) and kept on the engine -- a canvas
        # that spans 8 GPUs spends ~4 ms per rank in its one batch, so every small launch of the set-up shows in its scaling
        lo = r0 * rwidth + m if r0 > 0 else 0
        cache = self.engine.__dict__.setdefault('_crop_grid_cache', {})
        key = (H0, W0, m, r0, r1, self.patch)
        ent = cache.get(key)
        if ent is None:
            gy = torch.arange(r0, r1, dtype=torch.int32, device=dev) * rwidth
            gx = torch.arange(ncols, dtype=torch.int32, device=dev) * rwidth
            n_r = r1 - r0
            d_yx = torch.stack([gy[:, None].expand(n_r, ncols), gx[None, :].expand(n_r, ncols)], dim=2).reshape(-1, 2).contiguous()
            ys, xs = np.meshgrid(np.arange(r0, r1) * rwidth, np.arange(ncols) * rwidth, indexing='ij')
            crops_yx = np.stack([ys.ravel(), xs.ravel()], axis=1).astype(np.int32).reshape(-1, 2)
            d_local = d_yx.clone()
            d_local[:, 0] -= g0
            d_tiles = d_yx + m
            d_band = d_tiles.clone()
            d_band[:, 0] -= lo
            while len(cache) >= 16:
                cache.pop(next(iter(cache)))
            ent = cache[key] = (d_yx, crops_yx, d_local.contiguous(), d_tiles, d_band.contiguous())
        d_yx, self.crops_yx, self.d_crops_local, self.d_tiles_yx, self.d_tiles_yx_band = ent
        self._crops = None
        self.d_crops = d_yx
        self.tiles_yx = self.crops_yx + m
        # canvas rows owned by these crop rows under last-writer-wins (closed form): [r0*rw + m, r1*rw + m), the
        # first band starts at row 0 (rows < m are never written), the last one runs to the end of the canvas
        hi = r1 * rwidth + m if r1 < nrows else ph
        self.band = (lo, max(lo, hi))

    @property
    def crops(self) -> List[Tuple[int, int, int, int]]:
        """The reference's crop list [(y, x, h, w), ...] (style_transfer.py:33-48); built on first use."""
        if self._crops is None:
            self._crops = [(int(y), int(x), self.patch, self.patch) for y, x in self.crops_yx]
        return self._crops

    @property
    def geom(self) -> np.ndarray:
        """Host copy of the padded guidance [ph, pw, 1] (tests / debugging)."""
        return self.d_geom.cpu().numpy()[:, :, None]

    def gather(self, start: int, end: int) -> torch.Tensor:
        n = end - start
        out = torch.empty((n, 1, self.patch, self.patch), dtype=torch.float32, device=self.engine.device)
        crops = self.d_crops_local[start:end]                            # window-relative rows (no host-to-device scalar copies here)
        with torch.cuda.device(self.engine.device):
            _lib.call('nbe_gather_geom_patches', _lib.ptr(self.d_geom), int(self.d_geom.shape[0]), self.canvas_w,
                      _lib.ptr(crops.contiguous()), _lib.ptr(out), n, self.patch, _lib.stream())
        return out

    def place_band(self, band: torch.Tensor, tiles: torch.Tensor):
        """Place this job's tiles into ``band`` = canvas rows [band[0], band[1]) (a [rows, canvas_w, 4] uint8 tensor, e.g. a
        slice of the final canvas): ownership among the job's own tiles, clipped to the band -- exact in 'all' mode, where no
        other job's tile owns a pixel of these rows."""
        lo, hi = self.band
        rows = hi - lo
        n = len(self.crops_yx)
        if rows <= 0 or n == 0:
            return
        dev = self.engine.device
        yx = self.d_tiles_yx_band                                          # tile origins relative to the band (row-window jobs keep it ready)
        if yx is None:
            yx = self.d_tiles_yx.clone()
            yx[:, 0] -= lo
        owner = torch.empty((rows, self.canvas_w), dtype=torch.int32, device=dev)
        order = torch.arange(n, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.call('nbe_tile_owner_map', _lib.ptr(yx), n, self.tile, _lib.ptr(owner), rows, self.canvas_w, _lib.stream())
            _lib.call('nbe_place_tiles', _lib.ptr(tiles), _lib.ptr(yx), _lib.ptr(order), n, self.tile, _lib.ptr(owner), _lib.ptr(band),
                      rows, self.canvas_w, _lib.stream())

    def gather_indices(self, d_idx: torch.Tensor) -> torch.Tensor:
        """``gather`` for an arbitrary list of crop indices (int64 tensor on the device)."""
        n = int(d_idx.shape[0])
        crops = self.d_crops[d_idx].contiguous()
        out = torch.empty((n, 1, self.patch, self.patch), dtype=torch.float32, device=self.engine.device)
        with torch.cuda.device(self.engine.device):
            _lib.call('nbe_gather_geom_patches', _lib.ptr(self.d_geom), self.canvas_h, self.canvas_w, _lib.ptr(crops), _lib.ptr(out), n,
                      self.patch, _lib.stream())
        return out

    def owner_map(self) -> torch.Tensor:
        owner = torch.empty((self.canvas_h, self.canvas_w), dtype=torch.int32, device=self.engine.device)
        with torch.cuda.device(self.engine.device):
            _lib.call('nbe_tile_owner_map', _lib.ptr(self.d_tiles_yx), len(self.crops_yx), self.tile, _lib.ptr(owner),
                      self.canvas_h, self.canvas_w, _lib.stream())
        return owner

    def place(self, canvas: torch.Tensor, owner: torch.Tensor, tiles: torch.Tensor, start: int, end: int):
        order = torch.arange(start, end, dtype=torch.int32, device=self.engine.device)
        with torch.cuda.device(self.engine.device):
            _lib.call('nbe_place_tiles', _lib.ptr(tiles), _lib.ptr(self.d_tiles_yx[start:end]), _lib.ptr(order), end - start,
                      self.tile, _lib.ptr(owner), _lib.ptr(canvas), self.canvas_h, self.canvas_w, _lib.stream())

    def finish(self, canvas: torch.Tensor, on_white: bool, to_host: bool = True):
        """Crop back to the input size (paint_image_main.py:185-186); optional on-white composite (:179-183)."""
        m = self.crop_margin
        result = canvas[m:m + self.orig_shape[0], m:m + self.orig_shape[1], :]
        if not to_host:
            if on_white:
                alpha = result[..., 3:].to(torch.float32) / 255
                result = (result[..., :3].to(torch.float32) * alpha + 255 * (1 - alpha)).clip(0, 255).to(torch.uint8)
            return result
        # download through pinned memory (torch's caching host allocator makes the buffer free after the first call): ~3x
        # the bandwidth of a pageable copy, and the strided crop is compacted on the device first
        host = torch.empty(tuple(result.shape), dtype=result.dtype, pin_memory=True)
        host.copy_(result.contiguous(), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        result = host.numpy()
        if on_white:
            result = composite_on_white(result)
        return result


def _batch_opts(base: GanBrushOptions, z_per_patch: Optional[torch.Tensor], start: int, end: int, positions: torch.Tensor):
    o = GanBrushOptions()
    o.__dict__.update(base.__dict__)
    if z_per_patch is not None:
        o.style_z, o.style_ws = z_per_patch[start:end], None
    o.position = positions
    return o


def row_shards(nrows: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous crop-ROW ranges [r0, r1) per rank, as even as possible, earlier ranks take the extra rows
    (47 rows / 8 -> 6,6,6,6,6,6,6,5) -- the row view of ``shard_bounds`` for the dense grid."""
    base, extra = divmod(nrows, world_size)
    out = []
    for rank in range(world_size):
        r0 = rank * base + min(rank, extra)
        out.append((min(r0, nrows), min(r0 + base + (1 if rank < extra else 0), nrows)))
    return out


def exchange_bands(canvas: Optional[torch.Tensor], band: Optional[torch.Tensor], bands: Sequence[Tuple[int, int]], rank: int,
                   group=None) -> None:
    """The one inter-GPU step of a sharded canvas: every rank > 0 sends the canvas rows it owns (``band``, rows
    ``bands[rank]``) to rank 0, which receives each one straight into ``canvas[lo:hi]`` -- disjoint, contiguous row ranges,
    so there is no staging buffer and no placement pass on rank 0.  Device-agnostic (NCCL over NVLink on GPUs, gloo in the
    CPU tests); all transfers are posted as one batch."""
    import torch.distributed as dist
    ops = []
    if rank == 0:
        for r, (lo, hi) in enumerate(bands):
            if r != 0 and hi > lo:
                ops.append(dist.P2POp(dist.irecv, canvas[lo:hi], r, group))
    else:
        lo, hi = bands[rank]
        if hi > lo:
            ops.append(dist.P2POp(dist.isend, band, 0, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def gather_tiles(tiles_local: torch.Tensor, world: int, rank: int, group=None):
    """Sparse stitching modes: one gather of every rank's (equal-length, padded) tile buffer to rank 0 -> list of ``world``
    tensors on rank 0, None elsewhere.  Device-agnostic."""
    import torch.distributed as dist
    gathered = None
    if rank == 0:
        big = torch.empty((world,) + tuple(tiles_local.shape), dtype=tiles_local.dtype, device=tiles_local.device)
        gathered = list(big.unbind(0))
    dist.gather(tiles_local, gathered, dst=0, group=group)
    return gathered


def _plain_z_style(opts: GanBrushOptions) -> bool:
    """A brush the batch step's CUDA graph covers: styled by z, default colours, no UVS remapping, no custom noise."""
    return opts.style_ws is None and opts.style_z is not None and opts.color0 is None and opts.color1 is None \
        and opts.canvas_color is None and not opts.enable_uvs_mapping and not opts.custom_args


def _render_batch(engine, geom, opts, z_per_patch, z0, z1, pos, crop_margin, repeats):
    """One batch of a job -> uint8 tiles.  Batches whose size recurs (``repeats`` >= 2 in this job, or the engine has rendered a
    batch of that size before) replay the batch step's CUDA graph: one launch from the host instead of ~55; a size seen for
    the first time runs eagerly (capturing a graph costs more than one eager step)."""
    n = geom.shape[0]
    key = (n, int(crop_margin), engine.render_mode)
    hits = engine._batch_hits[key] = engine._batch_hits.get(key, 0) + 1          # how often this engine has seen the batch size
    if getattr(engine, 'use_batch_graph', False) and _plain_z_style(opts) and engine.G.flat_supported and engine.encoder.mode == 'bf16' \
            and (repeats >= 2 or hits >= 2 or key in engine._batch_sessions):
        z = z_per_patch[z0:z1] if z_per_patch is not None else opts.style_z
        z = z.to(engine.device, torch.float64)
        if z.shape[0] != n:
            z = z.expand(n, -1)
        return engine.render_tiles_graph(geom, z, pos, crop_margin)
    tiles, _ = engine.render_tiles(geom, _batch_opts(opts, z_per_patch, z0, z1, pos), crop_margin=crop_margin)
    return tiles


def _render_job_tiles(engine, job, opts, z_per_patch, z_offset, batch_size, crop_margin, out=None):
    """All crops of ``job`` through encoder -> generator -> composite in even batches -> uint8 tiles [n, T, T, 4]."""
    n = len(job.crops_yx)
    tiles_all = out if out is not None else torch.empty((n, job.tile, job.tile, 4), dtype=torch.uint8, device=engine.device)
    if n == 0:
        return tiles_all
    # even batches: a short tail batch costs almost a full launch sequence (276 patches -> 1 x 276, not 256 + 20)
    n_batches = max(1, -(-n // batch_size))
    if n <= batch_size * 5 // 4:
        n_batches = 1
    bs = max(1, -(-n // n_batches))
    for b0 in range(0, n, bs):
        b1 = min(b0 + bs, n)
        geom = job.gather(b0, b1)
        pos = job.d_crops[b0:b1].to(torch.int64)
        tiles_all[b0:b1] = _render_batch(engine, geom, opts, z_per_patch, z_offset + b0, z_offset + b1, pos, crop_margin, n // bs)
    return tiles_all


def stylize(engine: TriadPaintEngine, guidance: np.ndarray, opts: GanBrushOptions, crop_margin: int = 10,
            stitching_mode: str = 'all', feature_blending_level: int = 0, batch_size: int = 256, on_white: bool = False,
            z_per_patch: Optional[torch.Tensor] = None, group=None, return_job: bool = False, to_host: bool = True,
            distributed: bool = True, timings: Optional[dict] = None):
    """Stylize a whole guidance drawing.  guidance: [H,W,C] uint8 (last channel, 0 = stroke), host array or CUDA tensor;
    ``to_host=False`` returns the finished canvas as a CUDA tensor (device-resident in and out).

    With an initialised ``torch.distributed`` process group (one process per GPU) the crop rows are sharded across
    ranks and rank 0 returns the finished canvas (other ranks return None).  ``z_per_patch`` ([n_crops, z_dim])
    gives every patch its own style (style interpolation across the canvas, BASELINE config 5).
    ``distributed=False`` ignores the process group: this rank renders the whole canvas by itself.
    ``timings``: a dict that receives per-phase milliseconds (setup / render / place / exchange / finish); filling it
    synchronises the device between phases, so pass it on profiling runs only."""
    import time
    import torch.distributed as dist
    world, rank = 1, 0
    if distributed and dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = engine.device
    t_last = [time.perf_counter()]

    def mark(name):
        if timings is not None:
            torch.cuda.synchronize(dev)
            now = time.perf_counter()
            timings[name] = timings.get(name, 0.0) + (now - t_last[0]) * 1e3
            t_last[0] = now

    if feature_blending_level > 0:
        job = CanvasJob(engine, guidance, crop_margin, stitching_mode)
        if world > 1:
            # several GPUs: the layers before and after the blend point are sharded, the blend itself runs on rank 0
            # (_stylize_blended_phased); levels whose blend point the generator cannot be split at stay on one GPU
            res = engine.patch_width // 2 ** (feature_blending_level - 1)
            if not (_flat_blend_ok(engine) and _phased_blend_ok(engine, res, len(job.crops_yx))):
                raise RuntimeError('stylize: this feature-blending level / engine cannot be split at the blend point, so one canvas runs on '
                                   'one GPU (pass distributed=False and give every rank its own canvas)')
            canvas = _stylize_blended_phased(engine, job, opts, feature_blending_level, z_per_patch, batch_size, group, world, rank)
            if rank != 0:
                return (None, job) if return_job else None
        else:
            blend = _stylize_blended if os.environ.get('NBE_BLEND_SEQUENTIAL') else _stylize_blended_wavefront
            canvas = blend(engine, job, opts, feature_blending_level, z_per_patch)
        out = job.finish(canvas, on_white, to_host)
        return (out, job) if return_job else out

    if world > 1 and stitching_mode == 'all':
        # ---- dense grid on several GPUs: every rank owns a band of canvas rows (closed-form ownership), renders and places
        #      its own tiles there, and the bands travel to rank 0 in ONE batched send/recv straight into the final canvas
        m = int(crop_margin)
        H0, W0 = int(guidance.shape[0]), int(guidance.shape[1])
        nrows, ncols, rwidth, ph, pw = crop_grid(H0 + m, W0 + m, engine.patch_width, m * 2)
        shards = row_shards(nrows, world)
        r0, r1 = shards[rank]
        job = CanvasJob(engine, guidance, crop_margin, 'all', crop_rows=(r0, r1))
        bands = []
        for (a, b) in shards:
            lo = a * rwidth + m if a > 0 else 0
            hi = b * rwidth + m if b < nrows else ph
            bands.append((lo, max(lo, hi)))
        mark('setup')
        tiles_local = _render_job_tiles(engine, job, opts, z_per_patch, r0 * ncols, batch_size, crop_margin)
        mark('render')
        lo, hi = bands[rank]
        if rank == 0:
            canvas = torch.empty((ph, pw, 4), dtype=torch.uint8, device=dev)   # other ranks' rows are overwritten by the receive
            band = canvas[lo:hi]
        else:
            canvas = None
            band = torch.empty((hi - lo, pw, 4), dtype=torch.uint8, device=dev)
        band.zero_()
        job.place_band(band, tiles_local)
        mark('place')
        exchange_bands(canvas, band, bands, rank, group)
        mark('exchange')
        if rank != 0:
            return (None, job) if return_job else None
        out = job.finish(canvas, on_white, to_host)
        mark('finish')
        return (out, job) if return_job else out

    job = CanvasJob(engine, guidance, crop_margin, stitching_mode)
    bounds = shard_bounds(job.crops_yx, world)
    start, end = bounds[rank]
    max_n = max(e - s for s, e in bounds)
    mark('setup')
    # every rank's tile buffer has the same (maximum) length so that it can be gathered as it is
    tiles_local = torch.empty((max_n, job.tile, job.tile, 4), dtype=torch.uint8, device=dev)
    n_batches = max(1, -(-(end - start) // batch_size))
    if (end - start) <= batch_size * 5 // 4:
        n_batches = 1
    bs = max(1, -(-(end - start) // n_batches))
    for b0 in range(start, end, bs):
        b1 = min(b0 + bs, end)
        geom = job.gather(b0, b1)
        pos = job.d_crops[b0:b1].to(torch.int64)
        tiles_local[b0 - start:b1 - start] = _render_batch(engine, geom, opts, z_per_patch, b0, b1, pos, crop_margin, (end - start) // bs)
    mark('render')
    if world == 1:
        canvas = torch.zeros((job.canvas_h, job.canvas_w, 4), dtype=torch.uint8, device=dev)
        job.place(canvas, job.owner_map(), tiles_local, start, end)
        mark('place')
        out = job.finish(canvas, on_white, to_host)
        mark('finish')
        return (out, job) if return_job else out
    # ---- sparse crop lists on several GPUs: ownership is not a closed form, so finished tiles are gathered to rank 0
    #      (NCCL over NVLink) and placed there under the global ownership map
    gathered = gather_tiles(tiles_local, world, rank, group)
    mark('exchange')
    if rank != 0:
        return (None, job) if return_job else None
    canvas = torch.zeros((job.canvas_h, job.canvas_w, 4), dtype=torch.uint8, device=dev)
    owner = job.owner_map()
    for r, (s_, e_) in enumerate(bounds):
        if e_ > s_:
            job.place(canvas, owner, gathered[r], s_, e_)               # only the first e - s tiles of the buffer are read
    mark('place')
    out = job.finish(canvas, on_white, to_host)
    mark('finish')
    return (out, job) if return_job else out


# ------------------------------------------------------------------------------------------------ feature blending
def dirty_area_alpha(width: int, margin: int, crop_margin: int, device) -> torch.Tensor:
    """``PaintingHelper.generate_dirty_area_alpha`` for a full-patch dirty area (brush.py:159-187)."""
    lo = margin + crop_margin
    hi = lo + width - 2 * margin - 2 * crop_margin
    x = torch.linspace(0, width - 1, steps=width, device=device)
    gy, gx = torch.meshgrid(x, x, indexing='ij')
    dx = torch.min((gx - lo) ** 2, (gx - hi + 1) ** 2)
    dy = torch.min((gy - lo) ** 2, (gy - hi + 1) ** 2)
    d = dx + dy
    d[0:lo, lo:hi] = dy[0:lo, lo:hi]
    d[hi:, lo:hi] = dy[hi:, lo:hi]
    d[lo:hi, 0:lo] = dx[lo:hi, 0:lo]
    d[lo:hi, hi:] = dx[lo:hi, hi:]
    res = 1 - torch.sqrt(d) / margin
    res[res < 0] = 0
    res[lo:hi, lo:hi] = 1
    return res


class _Blended:
    def __init__(self, features, alpha):
        self.features, self.alpha = features, alpha


def _flat_blend_ok(engine: TriadPaintEngine) -> bool:
    """The flat tensor-core path can blend against a device-side feature canvas (generator.WindowBlend); FP32 engines and
    non-stock configurations keep the generic path (``NBE_BLEND_GENERIC`` forces it for A/B)."""
    return engine.G.flat_supported and engine.encoder.mode == 'bf16' and not engine.G._canvas_format \
        and os.environ.get('NBE_BLEND_GENERIC') is None


class _BlendContext:
    """Device-side state of feature-blended canvases of one shape: the NHWC bf16 feature canvas + its mask (re-zeroed per
    canvas), the dirty-area alpha, and one CUDA graph of the blended batch step per wavefront size (``engine.BatchSession``
    with ``blend=``), captured the second time a size is seen.  Kept on the engine, most recent shapes only."""

    def __init__(self, engine: TriadPaintEngine, fh: int, fw: int, res: int, margin: int, cm: int, crop_margin: int):
        dev = engine.device
        C = engine.G.cfg.channels(res)
        self.engine, self.res, self.cm, self.crop_margin = engine, res, cm, crop_margin
        self.base_alpha = dirty_area_alpha(res, margin, cm, dev).to(torch.float32).contiguous()
        self.fcanvas = torch.zeros((fh + res, fw + res, C), dtype=torch.bfloat16, device=dev)      # + res: windows never leave the buffer
        self.fmask = torch.zeros((fh + res, fw + res), dtype=torch.uint8, device=dev)
        self.sessions, self.hits = {}, {}

    def reset(self):
        self.fcanvas.zero_()
        self.fmask.zero_()

    def session(self, n: int):
        """The graph session for wavefronts of ``n`` patches, or None while the size has been seen only once."""
        from .engine import BatchSession
        sess = self.sessions.get(n)
        if sess is None:
            self.hits[n] = self.hits.get(n, 0) + 1
            if self.hits[n] < 2:
                return None
            sess = self.sessions[n] = BatchSession(self.engine, n, self.crop_margin, blend=dict(
                res=self.res, fcanvas=self.fcanvas, fmask=self.fmask, base_alpha=self.base_alpha, crop_margin=self.cm))
        return sess


def _blend_context(engine, fh, fw, res, margin, cm, crop_margin) -> _BlendContext:
    cache = engine.__dict__.setdefault('_blend_contexts', {})
    key = (fh, fw, res, margin, cm, crop_margin, engine.render_mode)
    ctx = cache.pop(key, None)
    if ctx is None:
        while len(cache) >= 2:                               # a 4096^2 level-2 feature canvas is 1.1 GB
            cache.pop(next(iter(cache)))
        ctx = _BlendContext(engine, fh, fw, res, margin, cm, crop_margin)
    cache[key] = ctx
    return ctx


def phased_shares(n: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous shares [a, b) of a list of ``n`` (wavefront-ordered) patches, one per rank."""
    return [(n * r // world, n * (r + 1) // world) for r in range(world)]


def exchange_shares(tensors: Sequence[torch.Tensor], shares: Sequence[Tuple[int, int]], rank: int, to_root: bool, group=None) -> None:
    """Move every rank's share of each tensor in ``tensors`` to rank 0 (``to_root``) or back, in ONE batched send/recv.  On
    rank 0 the tensors are full-length (its own share is already in place, share r lives at [a_r, b_r)); on rank r > 0 they hold
    that rank's share in their first b_r - a_r entries.  Device-agnostic (NCCL on GPUs, gloo in the CPU tests)."""
    import torch.distributed as dist
    world = len(shares)
    if world == 1:
        return
    peer = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
    ops = []
    for t in tensors:
        if rank == 0:
            for r in range(1, world):
                a, b = shares[r]
                if b > a:
                    ops.append(dist.P2POp(dist.irecv if to_root else dist.isend, t[a:b], peer(r), group))
        else:
            a, b = shares[rank]
            if b > a:
                ops.append(dist.P2POp(dist.isend if to_root else dist.irecv, t[:b - a], peer(0), group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


_RETURN_GROUPS = {}


def _return_group(group, world: int):
    """A second process group over the same ranks (its own NCCL communicator and stream), created once per main group: the
    blended feature maps travel back on it while the main communicator is still receiving.  Collective: every rank of ``group``
    reaches this call at the same point of ``_stylize_blended_phased``."""
    import torch.distributed as dist
    main = group if group is not None else dist.distributed_c10d._get_default_group()
    key = id(main)                                       # (a re-initialised default group is a new object: no stale communicator)
    g = _RETURN_GROUPS.get(key)
    if g is None:
        ranks = dist.get_process_group_ranks(group) if group is not None else list(range(world))
        g = _RETURN_GROUPS[key] = dist.new_group(ranks=ranks)
    return g


PHASED_BLEND_MAX_BYTES = 16 << 30          # budget for the per-canvas buffer of pre-blend feature maps (2.35 GB for 4096^2, level 2)


def _phased_blend_ok(engine: TriadPaintEngine, res: int, n_crops: int) -> bool:
    cfg = engine.G.cfg
    if os.environ.get('NBE_BLEND_WAVEFRONT_GRAPHS') is not None:                      # A/B switch: one whole forward per wavefront
        return False
    cbuf = cfg.channels(res) if res == cfg.img_resolution else cfg.block_in_channels(res * 2)
    return res <= cfg.img_resolution and res in cfg.block_resolutions and all(r <= res for r in cfg.geom_feature_resolutions) \
        and n_crops * res * (res + 1) * cbuf * 2 <= PHASED_BLEND_MAX_BYTES


def _stylize_blended_phased(engine: TriadPaintEngine, job: CanvasJob, opts: GanBrushOptions, level: int, z_per_patch, batch_size: int = 256,
                            group=None, world: int = 1, rank: int = 0):
    """Feature blending with the work split at the blend point.  Patch n's layers BEFORE the blend (encoder, mapping, blocks up to
    conv1 of the blended block) depend on nothing another patch computes, and its layers AFTER the blend only on its own blended
    feature map: the raster dependency of brush.py:190-242 runs through the blend alone (look up what earlier patches saved,
    blend, save the core).  So: (1) every patch's first half at full batch size, un-modulated feature maps into one buffer in
    wavefront order; (2) ``nbe_blend_window_nhwc_bf16`` wavefront by wavefront over slices of that buffer -- the only sequential
    part, ~140 launches of a few microseconds for a 4096^2 canvas; (3) every patch's second half at full batch size.  Same kernels
    on the same per-patch inputs as the wavefront schedule, hence the same bytes (every kernel is batch-invariant); 4096^2, level 2:
    the generator runs 2 x 9 batches of 256 instead of 141 of <= 24.

    Several GPUs (``world`` > 1, one process per GPU): phases (1) and (3) are sharded -- rank r takes the r-th contiguous share of
    the wavefront-ordered patch list -- and the blend runs on rank 0 over the gathered feature maps: the shares travel to rank 0,
    the ~140 blend launches run there (overlapped with the arrival of later shares and the return of finished ones, see below),
    the blended maps travel back, and the finished tiles are gathered on rank 0, which places them.  Every patch still goes through
    the same kernels in the same blend order: the canvas equals the single-GPU one bit for bit.  Returns the canvas on rank 0, None
    elsewhere."""
    import torch.distributed as dist
    from .generator import WindowBlend
    dev = engine.device
    down = 2 ** (level - 1)
    res = engine.patch_width // down
    C = engine.G.cfg.channels(res)                        # blended channels
    last = res == engine.G.cfg.img_resolution             # level 1: the blended map is the dense last feature map (no gap column, ToRGB follows)
    Cbuf = C if last else engine.G.cfg.block_in_channels(res * 2)   # channels of the buffer that feeds the next block (+ injected geometry at 32^2)
    pitch = res if last else res + 1
    fh, fw = int(math.ceil(job.canvas_h / down)), int(math.ceil(job.canvas_w / down))
    margin = 16 // down                                   # PaintingHelper.feature_blending_margin = 16
    cm = job.crop_margin // down
    snapped = (job.crops_yx // down) * down                                               # brush.py:253-258
    n_crops = len(job.crops_yx)
    waves = blending_wavefronts(job.crops_yx, engine.patch_width)
    d_order = torch.from_numpy(np.concatenate(waves)).to(dev)
    d_pos = job.d_crops.to(torch.int64)[d_order]
    d_cropsel = job.d_crops[d_order].contiguous()
    z_sel = z_per_patch.to(dev)[d_order] if z_per_patch is not None else None          # (a host tensor is uploaded once)
    shares = phased_shares(n_crops, world)
    s0, s1 = shares[rank]
    # rank 0 holds every patch's feature map (it blends them); the other ranks only their own share
    n_buf, b0 = (n_crops, 0) if rank == 0 else (s1 - s0, s0)
    key = (n_buf, res, Cbuf)
    buf = engine.__dict__.get('_phased_blend_buf')
    if buf is None or buf[0] != key:
        buf = engine.__dict__['_phased_blend_buf'] = (key, torch.zeros((n_buf, res, pitch, Cbuf), dtype=torch.bfloat16, device=dev),
                                                      torch.empty((n_buf, C), dtype=torch.float32, device=dev))
    _, X, NS = buf                                         # (the gap column of X is never written: zero from the allocation on)
    tiles_all = torch.empty((n_buf, job.tile, job.tile, 4), dtype=torch.uint8, device=dev)

    def chunk_opts(sl):
        o = GanBrushOptions()
        o.__dict__.update(opts.__dict__)
        if z_sel is not None:
            o.style_z, o.style_ws = z_sel[sl], None
        o.position = d_pos[sl]
        return o

    def exchange(tensors, to_root: bool):
        exchange_shares(tensors, shares, rank, to_root, group)

    # even batches, like _render_job_tiles: a short tail batch costs almost a full launch sequence (276 patches -> 1 x 276)
    n_own = s1 - s0
    n_batches = 1 if n_own <= batch_size * 5 // 4 else -(-n_own // batch_size)
    batch_size = max(1, -(-n_own // n_batches)) if n_own else batch_size
    with torch.cuda.device(dev):
        for c0 in range(s0, s1, batch_size):                                              # (1) everything before the blend
            sl = slice(c0, min(s1, c0 + batch_size))
            loc = slice(sl.start - b0, sl.stop - b0)
            n = sl.stop - sl.start
            geom = torch.empty((n, 1, job.patch, job.patch), dtype=torch.float32, device=dev)
            _lib.call('nbe_gather_geom_patches', _lib.ptr(job.d_geom), job.canvas_h, job.canvas_w, _lib.ptr(d_cropsel[sl]), _lib.ptr(geom), n,
                      job.patch, _lib.stream())
            ns_ = engine.render_split_pre(geom, chunk_opts(sl), res, X[loc])
            if ns_ is not None:
                NS[loc] = ns_
        # (2) the blend, in dependency order.  Several GPUs: the shares arrive one after the other (per-share receives in rank order on
        # the main communicator) while rank 0 already blends the wavefronts whose patches are complete, and a share travels back -- on a
        # SECOND communicator, so that returns are not queued behind the receives -- as soon as the last wavefront that touches it is done:
        # of the 2 x 2 GB through rank 0's links only the first share in and the last share out are not hidden behind the blend.
        # (a sub-group of the world keeps the two batched exchanges: creating the second communicator is collective over ALL ranks)
        overlap = world > 1 and dev.type == 'cuda' and group is None and os.environ.get('NBE_BLEND_NO_EXCHANGE_OVERLAP') is None
        back = _return_group(group, world) if overlap else None
        peer = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
        feats = [X] if last else [X, NS]
        if not overlap:
            exchange(feats, to_root=True)
        elif rank != 0 and s1 > s0:
            out_reqs = [dist.isend(t[:s1 - s0], peer(0), group) for t in feats]
        if rank == 0:
            recv = {}
            if overlap:
                for r in range(1, world):
                    a, b = shares[r]
                    recv[r] = [dist.irecv(t[a:b], peer(r), group) for t in feats] if b > a else []
            ctx = _blend_context(engine, fh, fw, res, margin, cm, job.crop_margin)
            ctx.reset()
            d_fyx = torch.from_numpy(np.ascontiguousarray(snapped // down).astype(np.int32)).to(dev)[d_order].contiguous()
            off, arrived, returned, sends = 0, 0, 0, []
            for idx in waves:
                sl = slice(off, off + len(idx))
                off += len(idx)
                while overlap and arrived + 1 < world and shares[arrived + 1][0] < off:    # this wavefront reaches into the next share(s)
                    arrived += 1
                    for q in recv[arrived]:
                        q.wait()                                                          # (the compute stream waits, not the host)
                WindowBlend(res, ctx.fcanvas, ctx.fmask, d_fyx[sl], ctx.base_alpha, cm).apply(X[sl], pitch, C, None if last else NS[sl], len(idx))
                while overlap and returned + 1 < world and shares[returned + 1][1] <= off:  # share complete: send it home
                    returned += 1
                    a, b = shares[returned]
                    if b > a:
                        sends.append(dist.isend(X[a:b], peer(returned), back))
            for q in sends:
                q.wait()
        elif overlap and s1 > s0:
            dist.irecv(X[:s1 - s0], peer(0), back).wait()
            for q in out_reqs:
                q.wait()
        if not overlap:
            exchange([X], to_root=False)
        for c0 in range(s0, s1, batch_size):                                              # (3) everything after it
            sl = slice(c0, min(s1, c0 + batch_size))
            loc = slice(sl.start - b0, sl.stop - b0)
            tiles_all[loc] = engine.render_split_post(X[loc], chunk_opts(sl), res, crop_margin=job.crop_margin)
        exchange([tiles_all], to_root=True)
        if rank != 0:
            return None
        canvas = torch.zeros((job.canvas_h, job.canvas_w, 4), dtype=torch.uint8, device=dev)
        owner = torch.empty((job.canvas_h, job.canvas_w), dtype=torch.int32, device=dev)
        # last writer wins in RASTER order: tile k of the wavefront-ordered list is crop d_order[k]
        raster_yx = torch.from_numpy(np.ascontiguousarray(snapped + job.crop_margin).astype(np.int32)).to(dev)
        _lib.call('nbe_tile_owner_map', _lib.ptr(raster_yx), n_crops, job.tile, _lib.ptr(owner), job.canvas_h, job.canvas_w, _lib.stream())
        raster_tiles = torch.empty_like(tiles_all)
        raster_tiles[d_order] = tiles_all
        order = torch.arange(n_crops, dtype=torch.int32, device=dev)
        _lib.call('nbe_place_tiles', _lib.ptr(raster_tiles), _lib.ptr(raster_yx), _lib.ptr(order), n_crops, job.tile, _lib.ptr(owner),
                  _lib.ptr(canvas), job.canvas_h, job.canvas_w, _lib.stream())
    return canvas


def _stylize_blended_flat(engine: TriadPaintEngine, job: CanvasJob, opts: GanBrushOptions, level: int, z_per_patch, sequential: bool):
    """Feature blending on the flat bf16 path: the feature canvas is a device-side NHWC bf16 tensor, and look-up, blend,
    core write-back and the next layer's modulation of a whole batch of windows are ONE kernel inside the generator
    (``generator.WindowBlend`` -> ``nbe_blend_window_nhwc_bf16``) instead of indexed torch copies around it.
    ``sequential``: one patch per generator call in raster order (the reference's loop, brush.py:190-242); otherwise one
    anti-diagonal wavefront of mutually independent patches per call (``blending_wavefronts``) -- the two give the same bytes
    because every kernel on the path is batch-invariant.  Wavefront sizes that recur (the second canvas of a shape onwards)
    replay one CUDA graph each: gather + three small copies + one graph launch per wavefront."""
    from .generator import WindowBlend
    dev = engine.device
    down = 2 ** (level - 1)
    res = engine.patch_width // down
    fh, fw = int(math.ceil(job.canvas_h / down)), int(math.ceil(job.canvas_w / down))
    margin = 16 // down                                   # PaintingHelper.feature_blending_margin = 16
    cm = job.crop_margin // down
    ctx = _blend_context(engine, fh, fw, res, margin, cm, job.crop_margin)
    ctx.reset()
    snapped = (job.crops_yx // down) * down                                               # brush.py:253-258
    tiles_yx = torch.from_numpy(np.ascontiguousarray(snapped + job.crop_margin).astype(np.int32)).to(dev)
    n_crops = len(job.crops_yx)
    tiles_all = torch.empty((n_crops, job.tile, job.tile, 4), dtype=torch.uint8, device=dev)
    waves = [np.array([i], dtype=np.int64) for i in range(n_crops)] if sequential else blending_wavefronts(job.crops_yx, engine.patch_width)
    # one upload for the whole schedule; the loop below never synchronises with the host
    d_order = torch.from_numpy(np.concatenate(waves)).to(dev)
    d_fyx = torch.from_numpy(np.ascontiguousarray(snapped // down).astype(np.int32)).to(dev)[d_order].contiguous()
    d_pos = job.d_crops.to(torch.int64)[d_order]
    d_cropsel = job.d_crops[d_order].contiguous()
    z_sel = z_per_patch.to(dev)[d_order].to(torch.float64) if z_per_patch is not None else None
    graphs = getattr(engine, 'use_batch_graph', False) and _plain_z_style(opts) and not sequential
    z_one = opts.style_z.to(dev, torch.float64) if graphs and z_sel is None else None
    off = 0
    with torch.cuda.device(dev):
        for idx in waves:
            n = len(idx)
            sl = slice(off, off + n)
            off += n
            sess = ctx.session(n) if graphs else None
            if sess is not None:
                _lib.call('nbe_gather_geom_patches', _lib.ptr(job.d_geom), job.canvas_h, job.canvas_w, _lib.ptr(d_cropsel[sl]),
                          _lib.ptr(sess._geom), n, job.patch, _lib.stream())
                sess._z.copy_(z_sel[sl] if z_sel is not None else (z_one if z_one.shape[0] == n else z_one.expand(n, -1)), non_blocking=True)
                sess._pos.copy_(d_pos[sl], non_blocking=True)
                sess._fyx.copy_(d_fyx[sl], non_blocking=True)
                sess._graph.replay()
                tiles = sess._out
            else:
                geom = torch.empty((n, 1, job.patch, job.patch), dtype=torch.float32, device=dev)
                _lib.call('nbe_gather_geom_patches', _lib.ptr(job.d_geom), job.canvas_h, job.canvas_w, _lib.ptr(d_cropsel[sl]), _lib.ptr(geom), n,
                          job.patch, _lib.stream())
                o = GanBrushOptions()
                o.__dict__.update(opts.__dict__)
                if z_sel is not None:
                    o.style_z, o.style_ws = z_sel[sl], None
                o.position = d_pos[sl]
                wb = WindowBlend(res, ctx.fcanvas, ctx.fmask, d_fyx[sl], ctx.base_alpha, cm)
                tiles, _ = engine.render_tiles(geom, o, crop_margin=job.crop_margin, window_blend=wb)
            tiles_all[d_order[sl]] = tiles
        canvas = torch.zeros((job.canvas_h, job.canvas_w, 4), dtype=torch.uint8, device=dev)
        owner = torch.empty((job.canvas_h, job.canvas_w), dtype=torch.int32, device=dev)
        _lib.call('nbe_tile_owner_map', _lib.ptr(tiles_yx), n_crops, job.tile, _lib.ptr(owner), job.canvas_h, job.canvas_w, _lib.stream())
        order = torch.arange(n_crops, dtype=torch.int32, device=dev)
        _lib.call('nbe_place_tiles', _lib.ptr(tiles_all), _lib.ptr(tiles_yx), _lib.ptr(order), n_crops, job.tile, _lib.ptr(owner),
                  _lib.ptr(canvas), job.canvas_h, job.canvas_w, _lib.stream())
    return canvas


class RasterFeatureCanvas:
    """``FeatureCanvas`` + ``PaintingHelper._get_blended_features`` / ``update_blended_features`` (brush.py:33-92, 190-242) on
    the generic generator path (``return_features`` / ``blended_features``; FP32 engines and non-stock configurations): one
    patch at a time, in the order the caller renders them.  The flat bf16 path keeps its feature canvas in NHWC bf16 and
    blends inside the generator instead (``generator.WindowBlend``)."""

    def __init__(self, engine: TriadPaintEngine, level: int, fh: int, fw: int, margin: int = 16):
        self.engine = engine
        self.down = 2 ** (level - 1)
        self.res = engine.patch_width // self.down
        self.fh, self.fw = int(fh), int(fw)
        self.margin = margin // self.down                     # PaintingHelper.feature_blending_margin = 16
        self.features, self.mask = None, None                 # allocated by the first patch (brush.py:48-57)
        self._base = {}

    def _alpha(self, cm: int):
        b = self._base.get(cm)
        if b is None:
            a = dirty_area_alpha(self.res, self.margin, cm, self.engine.device)
            b = self._base[cm] = (a, a > 0.99)
        return b

    def render(self, geom: torch.Tensor, opts: GanBrushOptions, y: int, x: int, crop_margin: int) -> torch.Tensor:
        """Render the patch whose window starts at canvas (y, x) (already snapped to the feature grid), blended with what
        earlier patches saved, and save its core -> uint8 tiles [1, T, T, 4] on the device."""
        res, down = self.res, self.down
        cm = crop_margin // down
        ys, xs = y // down, x // down
        alpha, update = self._alpha(cm)
        blended = {}
        if self.mask is not None:
            m = self.mask[ys:ys + res, xs:xs + res]
            update = update | (m & (alpha > 0))
            alpha = alpha.clone()
            alpha[~m] = 1
            blended = {res: _Blended(self.features[..., ys:ys + res, xs:xs + res], (1 - alpha)[None, None])}
        if cm > 0:
            update = update.clone()
            update[:cm, :] = False
            update[-cm:, :] = False
            update[:, :cm] = False
            update[:, -cm:] = False
        tiles, raw = self.engine.render_tiles(geom, opts, crop_margin=crop_margin, return_features=[res], blended_features=blended)
        feat = raw[f'features{res}']
        if self.features is None:
            # + res: a window that starts inside the canvas never leaves the buffer
            self.features = torch.zeros((1, feat.shape[1], self.fh + res, self.fw + res), dtype=feat.dtype, device=feat.device)
            self.mask = torch.zeros((self.fh + res, self.fw + res), dtype=torch.bool, device=feat.device)
        self.mask[ys:ys + res, xs:xs + res][update] = True
        um = update[None, None].expand(-1, feat.shape[1], -1, -1)
        self.features[..., ys:ys + res, xs:xs + res][um] = feat[um]
        return tiles


def _stylize_blended(engine: TriadPaintEngine, job: CanvasJob, opts: GanBrushOptions, level: int, z_per_patch):
    """Raster-order execution with a feature canvas (brush.py:33-92, 190-242): patch n blends the features saved
    by earlier overlapping patches into its own post-b(128/2^(level-1)) activations, then saves its core."""
    if _flat_blend_ok(engine):
        return _stylize_blended_flat(engine, job, opts, level, z_per_patch, sequential=True)
    dev = engine.device
    down = 2 ** (level - 1)
    fh, fw = int(math.ceil(job.canvas_h / down)), int(math.ceil(job.canvas_w / down))
    fc = RasterFeatureCanvas(engine, level, fh, fw)
    canvas = torch.zeros((job.canvas_h, job.canvas_w, 4), dtype=torch.uint8, device=dev)
    for i, (y, x, _, _) in enumerate(job.crops):
        ys, xs = y // down * down, x // down * down                           # snap to the feature grid (brush.py:253-258)
        geom = job.gather(i, i + 1)
        pos = job.d_crops[i:i + 1].to(torch.int64)
        tiles = fc.render(geom, _batch_opts(opts, z_per_patch, i, i + 1, pos), ys, xs, job.crop_margin)
        ty, tx = ys + job.crop_margin, xs + job.crop_margin
        canvas[ty:ty + job.tile, tx:tx + job.tile] = tiles[0]                 # raster order = last writer wins
    return canvas


def blending_wavefronts(crops_yx: np.ndarray, patch: int):
    """Group raster-ordered crops into wavefronts of mutually independent patches.

    Patch n depends on the earlier patches whose 128-px windows overlap its own (it reads the features they saved) and
    must run before the later ones that overlap it.  On the regular crop grid (stride = patch - 2 * margin > patch / 2)
    a window only overlaps its 8 grid neighbours, so with grid coordinates (r, c) the patches of one anti-diagonal
    ``2 r + c = const`` neither read nor write each other's windows, and every raster predecessor among the neighbours --
    (r, c-1), (r-1, c-1), (r-1, c), (r-1, c+1) -- lies on an earlier one.  Executing the diagonals in order, each as one
    batch, therefore reproduces the sequential raster loop of brush.py:190-242 exactly.
    Returns a list of int64 index arrays (raster indices, ascending inside a wavefront)."""
    ys = np.unique(crops_yx[:, 0])
    xs = np.unique(crops_yx[:, 1])
    if (len(ys) > 1 and np.diff(ys).min() * 2 <= patch) or (len(xs) > 1 and np.diff(xs).min() * 2 <= patch):
        raise RuntimeError('blending_wavefronts: crop stride must exceed half a patch (windows overlap more than their neighbours)')
    r = np.searchsorted(ys, crops_yx[:, 0])
    c = np.searchsorted(xs, crops_yx[:, 1])
    w = 2 * r + c
    order = np.argsort(w, kind='stable')
    bounds = np.flatnonzero(np.diff(w[order])) + 1
    return [g.astype(np.int64) for g in np.split(order, bounds)]


def _stylize_blended_wavefront(engine: TriadPaintEngine, job: CanvasJob, opts: GanBrushOptions, level: int, z_per_patch):
    """Feature blending with the raster semantics of ``_stylize_blended`` (brush.py:33-92, 190-242), executed one
    anti-diagonal wavefront of independent patches at a time (see ``blending_wavefronts``): ~2 rows + cols batched
    generator calls instead of rows x cols single-patch ones.  The feature canvas starts as zeros with an all-False
    mask (alpha = 1 wherever nothing was saved yet)."""
    if _flat_blend_ok(engine):
        if _phased_blend_ok(engine, engine.patch_width // 2 ** (level - 1), len(job.crops_yx)):
            return _stylize_blended_phased(engine, job, opts, level, z_per_patch)
        return _stylize_blended_flat(engine, job, opts, level, z_per_patch, sequential=False)
    dev = engine.device
    down = 2 ** (level - 1)
    res = engine.patch_width // down
    fh, fw = int(math.ceil(job.canvas_h / down)), int(math.ceil(job.canvas_w / down))
    margin = 16 // down                                   # PaintingHelper.feature_blending_margin = 16
    cm = job.crop_margin // down
    base_alpha = dirty_area_alpha(res, margin, cm, dev)
    base_update = base_alpha > 0.99
    inner = torch.ones((res, res), dtype=torch.bool, device=dev)
    if cm > 0:
        inner[:cm, :] = False
        inner[-cm:, :] = False
        inner[:, :cm] = False
        inner[:, -cm:] = False
    C = engine.G.cfg.channels(res)
    features = torch.zeros((C, fh + res, fw + res), dtype=torch.float32, device=dev)      # + res: windows never leave the buffer
    mask = torch.zeros((fh + res, fw + res), dtype=torch.bool, device=dev)
    snapped = (job.crops_yx // down) * down                                               # brush.py:253-258
    tiles_yx = torch.from_numpy(np.ascontiguousarray(snapped + job.crop_margin).astype(np.int32)).to(dev)
    n_crops = len(job.crops_yx)
    tiles_all = torch.empty((n_crops, job.tile, job.tile, 4), dtype=torch.uint8, device=dev)
    ar = torch.arange(res, device=dev)
    waves = []
    for idx in blending_wavefronts(job.crops_yx, engine.patch_width):
        # the raster loop does not blend its very first patch (no feature canvas yet, brush.py:196-201): on the bf16 path
        # "blend with alpha 0" and "no blend" round differently (the un-blended layer fuses the next layer's modulation),
        # so patch 0 runs on its own, un-blended, exactly as in the sequential loop
        if idx[0] == 0 and len(idx) > 1:
            waves += [idx[:1], idx[1:]]
        else:
            waves.append(idx)
    # one upload for the whole schedule; the loop below never synchronises with the host
    d_order = torch.from_numpy(np.concatenate(waves)).to(dev)
    d_fy = torch.from_numpy(np.ascontiguousarray(snapped[:, 0] // down)).to(dev)[d_order]
    d_fx = torch.from_numpy(np.ascontiguousarray(snapped[:, 1] // down)).to(dev)[d_order]
    d_pos = job.d_crops.to(torch.int64)[d_order]
    d_cropsel = job.d_crops[d_order].contiguous()
    z_sel = z_per_patch[d_order] if z_per_patch is not None else None
    alpha_pos = (base_alpha > 0)[None]
    one = torch.ones((), device=dev)
    # feature-canvas row / column of every window element of every patch, once for the whole schedule (views below)
    all_rows = (d_fy[:, None] + ar[None, :])[:, :, None].expand(-1, res, res)
    all_cols = (d_fx[:, None] + ar[None, :])[:, None, :].expand(-1, res, res)
    base_alpha_b = base_alpha[None]
    off = 0
    for idx in waves:
        n = len(idx)
        sl = slice(off, off + n)
        off += n
        rows, cols = all_rows[sl], all_cols[sl]                                           # [n,res,res]
        m = mask[rows, cols]                                                              # [n,res,res]
        update = (base_update[None] | (m & alpha_pos)) & inner[None]
        alpha = torch.where(m, base_alpha_b, one)
        saved = features[:, rows, cols].permute(1, 0, 2, 3)                               # [n,C,res,res]
        geom = torch.empty((n, 1, job.patch, job.patch), dtype=torch.float32, device=dev)
        _lib.call('nbe_gather_geom_patches', _lib.ptr(job.d_geom), job.canvas_h, job.canvas_w, _lib.ptr(d_cropsel[sl]), _lib.ptr(geom), n,
                  job.patch, _lib.stream())
        o = GanBrushOptions()
        o.__dict__.update(opts.__dict__)
        if z_sel is not None:
            o.style_z, o.style_ws = z_sel[sl], None
        o.position = d_pos[sl]
        blended = {} if idx[0] == 0 else {res: _Blended(saved, (1 - alpha)[:, None])}
        tiles, raw = engine.render_tiles(geom, o, crop_margin=job.crop_margin, return_features=[res], blended_features=blended)
        feat = raw[f'features{res}']                                                      # [n,C,res,res]
        # windows of one wavefront are disjoint: whole windows are written back (new value where `update`, old elsewhere),
        # which needs no data-dependent index list and therefore no host synchronisation
        mask[rows, cols] = m | update
        features[:, rows, cols] = torch.where(update[:, None], feat, saved).permute(1, 0, 2, 3)
        tiles_all[d_order[sl]] = tiles
    canvas = torch.zeros((job.canvas_h, job.canvas_w, 4), dtype=torch.uint8, device=dev)
    owner = torch.empty((job.canvas_h, job.canvas_w), dtype=torch.int32, device=dev)
    _lib.call('nbe_tile_owner_map', _lib.ptr(tiles_yx), n_crops, job.tile, _lib.ptr(owner), job.canvas_h, job.canvas_w, _lib.stream())
    order = torch.arange(n_crops, dtype=torch.int32, device=dev)
    _lib.call('nbe_place_tiles', _lib.ptr(tiles_all), _lib.ptr(tiles_yx), _lib.ptr(order), n_crops, job.tile, _lib.ptr(owner),
              _lib.ptr(canvas), job.canvas_h, job.canvas_w, _lib.stream())
    return canvas
