"""CPU: host-side logic of the patch scheduler -- crop grid, sharding, and the N>1 gather path on gloo (world 2)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import REPO
from oracle import neube_oracle as O
from brushstroke_engine_b200 import stylizer, synthetic, params as P


def test_crops_match_oracle_and_reference_sizes():
    for (h, w, mode) in ((300, 260, 'all'), (2000, 2000, 'all'), (500, 333, 'full'), (87, 89, 'all')):
        guidance = synthetic.synthetic_guidance(h, w, num_lines=5, seed=h, radii=(3, 9))
        a, pa = stylizer.generate_stitching_crops(stylizer.pad_geo(guidance, 10), 128, mode, 20)
        b, pb = O.generate_stitching_crops(O.pad_geo(guidance, 10), 128, mode, 20)
        assert a == b and np.array_equal(pa, pb)
    crops, padded = stylizer.generate_stitching_crops(stylizer.pad_geo(np.full((4096, 4096, 1), 255, np.uint8), 10), 128, 'all', 20)
    assert len(crops) == 2209 and padded.shape[:2] == (4264, 4264)


@pytest.mark.parametrize('world', [1, 2, 3, 4, 8, 64])
def test_shard_crops_partitions_rows(world):
    crops, _ = stylizer.generate_stitching_crops(stylizer.pad_geo(np.full((4096, 4096, 1), 255, np.uint8), 10), 128, 'all', 20)
    bounds = [stylizer.shard_crops(crops, world, r) for r in range(world)]
    assert bounds[0][0] == 0 and bounds[-1][1] == len(crops)
    for (s0, e0), (s1, e1) in zip(bounds[:-1], bounds[1:]):
        assert e0 == s1                                    # contiguous, in rank order (keeps last-writer-wins across seams)
    for s, e in bounds:
        if e > s:
            assert crops[s][1] == 0                        # bands start at a row boundary
    if world == 8:
        assert [(e - s) // 47 for s, e in bounds] == [6, 6, 6, 6, 6, 6, 6, 5]
    # ragged crop lists (mode != 'all') still partition
    guidance = synthetic.synthetic_guidance(900, 700, num_lines=4, seed=1, radii=(3,))
    rag, _ = stylizer.generate_stitching_crops(stylizer.pad_geo(guidance, 10), 128, 'full', 20)
    b2 = [stylizer.shard_crops(rag, world, r) for r in range(world)]
    assert sum(e - s for s, e in b2) == len(rag)


def test_style_seed_conventions():
    z = P.style_z_from_seed(594)
    assert z.dtype == torch.float64 and z.shape == (1, 64)
    assert np.allclose(z.numpy(), np.random.RandomState(594).randn(1, 64))
    zi = P.interpolated_style_z(1, 2, 0.25)
    assert np.allclose(zi.numpy(), 0.25 * np.random.RandomState(1).randn(1, 64) + 0.75 * np.random.RandomState(2).randn(1, 64))
    assert P.GeneratorConfig().num_ws == 12
    assert [P.GeneratorConfig().block_in_channels(r) for r in (8, 16, 32, 64, 128)] == [128, 128, 144, 384, 128]


def _gloo_worker(rank, world, port, tmp):
    """Runs the scheduler's own exchange functions (stylizer.row_shards / exchange_bands / shard_bounds / gather_tiles -- the
    ones stylize() calls on NCCL) on gloo with CPU tensors; 'rendered' tiles are fakes whose pixels carry the tile's raster
    index, placed on the CPU by the oracle's raster loop, so every placement or ownership error is visible."""
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sys.path.insert(0, REPO)
    m, T = 10, 108
    guidance = synthetic.synthetic_guidance(300, 260, num_lines=10, seed=5, radii=(1, 3, 9))
    crops, padded = stylizer.generate_stitching_crops(stylizer.pad_geo(guidance, m), 128, 'all', 2 * m)
    ph, pw = padded.shape[:2]
    metas = [(c[0] + m, c[1] + m) for c in crops]
    fake = lambda i: np.full((T, T, 4), i % 251, np.uint8)
    ref = O.place_tiles((ph, pw), [fake(i) for i in range(len(crops))], metas)
    ok = {}
    # ---- dense grid: bands of owned canvas rows, one batched send/recv into the final canvas
    nrows, ncols, rwidth, ph2, pw2 = stylizer.crop_grid(300 + m, 260 + m, 128, 2 * m)
    assert (ph2, pw2) == (ph, pw) and nrows * ncols == len(crops)
    shards = stylizer.row_shards(nrows, world)
    bands = []
    for (a, b) in shards:
        lo = a * rwidth + m if a > 0 else 0
        hi = b * rwidth + m if b < nrows else ph
        bands.append((lo, max(lo, hi)))
    assert bands[0][0] == 0 and bands[-1][1] == ph and all(x[1] == y[0] for x, y in zip(bands[:-1], bands[1:]))
    r0, r1 = shards[rank]
    mine = list(range(r0 * ncols, r1 * ncols))
    own = O.place_tiles((ph, pw), [fake(i) for i in mine], [metas[i] for i in mine])     # this rank's tiles only
    lo, hi = bands[rank]
    canvas = torch.full((ph, pw, 4), 77, dtype=torch.uint8) if rank == 0 else None
    if rank == 0:
        canvas[lo:hi] = torch.from_numpy(own[lo:hi])
        band = canvas[lo:hi]
    else:
        band = torch.from_numpy(np.ascontiguousarray(own[lo:hi]))
    stylizer.exchange_bands(canvas, band, bands, rank)
    if rank == 0:
        ok['bands'] = bool(np.array_equal(canvas.numpy(), ref))
    # ---- sparse lists: padded tile buffers gathered to rank 0, placed under global raster order
    yx = np.array([(c[0], c[1]) for c in crops], dtype=np.int32)
    bounds = stylizer.shard_bounds(yx, world)
    assert bounds == [stylizer.shard_crops(crops, world, r) for r in range(world)]
    s, e = bounds[rank]
    max_n = max(b - a for a, b in bounds)
    padded_t = torch.zeros((max_n, T, T, 4), dtype=torch.uint8)
    for k, i in enumerate(range(s, e)):
        padded_t[k] = torch.from_numpy(fake(i))
    gathered = stylizer.gather_tiles(padded_t, world, rank)
    if rank == 0:
        all_tiles = torch.cat([gathered[r][: b - a] for r, (a, b) in enumerate(bounds)]).numpy()
        ok['tiles'] = bool(np.array_equal(O.place_tiles((ph, pw), list(all_tiles), metas), ref))
        np.save(os.path.join(tmp, 'ok.npy'), np.array([ok['bands'], ok['tiles']]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_multi_rank_exchange_reassembles_canvas(tmp_path, world):
    """world_size 2 and 3 on gloo: row sharding + band exchange (dense grid) and tile gather (sparse lists) both reproduce
    the single-process raster-loop canvas bit for bit."""
    port = 29500 + (os.getpid() % 500) + world
    mp.spawn(_gloo_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = np.load(os.path.join(str(tmp_path), 'ok.npy'))
    assert bool(res[0]) and bool(res[1])


def _gloo_share_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    n = 11                                                     # 11 patches over 2 / 3 / 4 ranks: uneven shares, an empty one never
    shares = stylizer.phased_shares(n, world)
    a, b = shares[rank]
    full = torch.arange(n * 6, dtype=torch.float32).reshape(n, 2, 3)
    aux = torch.arange(n, dtype=torch.int64) * 10
    if rank == 0:
        x, y = torch.full_like(full, -1.0), torch.full_like(aux, -1)
        x[a:b], y[a:b] = full[a:b], aux[a:b]
    else:
        x, y = full[a:b].clone(), aux[a:b].clone()
    stylizer.exchange_shares([x, y], shares, rank, to_root=True)                     # what the phased feature blending gathers
    ok = True
    if rank == 0:
        ok = bool(torch.equal(x, full) and torch.equal(y, aux))
        x = x * 2                                                                    # "blend" on rank 0 ...
    else:
        x = torch.zeros_like(x)
    stylizer.exchange_shares([x], shares, rank, to_root=False)                       # ... and every share travels back
    ok = ok and bool(torch.equal(x[a:b] if rank == 0 else x, full[a:b] * 2))
    res = torch.tensor([1 if ok else 0])
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    if rank == 0:
        np.save(os.path.join(tmp, 'ok_shares.npy'), np.array([int(res[0])]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_phased_blend_share_exchange(tmp_path, world):
    """world_size 2 and 3 on gloo: the share exchange of the phased feature blending (stylizer.exchange_shares -- every rank's
    contiguous share of the wavefront-ordered patch list to rank 0 and back) reassembles and returns the data exactly."""
    assert stylizer.phased_shares(2209, 8)[0] == (0, 276) and stylizer.phased_shares(2209, 8)[-1][1] == 2209
    assert all(b >= a for a, b in stylizer.phased_shares(3, 8)) and sum(b - a for a, b in stylizer.phased_shares(3, 8)) == 3
    port = 29000 + (os.getpid() % 400) + world
    mp.spawn(_gloo_share_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert int(np.load(os.path.join(str(tmp_path), 'ok_shares.npy'))[0]) == 1


def test_phased_blend_eligibility():
    """Which feature-blending levels the stylizer splits at the blend point (stylizer._phased_blend_ok): a block of the generator at
    or above every geometry injection, with the per-canvas buffer of pre-blend feature maps inside its memory budget."""
    import types
    from brushstroke_engine_b200 import params as P
    cfg = P.GeneratorConfig()
    eng = types.SimpleNamespace(G=types.SimpleNamespace(cfg=cfg))
    assert stylizer._phased_blend_ok(eng, 64, 2209) and stylizer._phased_blend_ok(eng, 32, 2209) and stylizer._phased_blend_ok(eng, 128, 2209)
    assert not stylizer._phased_blend_ok(eng, 16, 2209)                      # below the 32^2 injection: the encoder still has to write into it
    assert not stylizer._phased_blend_ok(eng, 48, 2209)                      # not a block resolution
    assert not stylizer._phased_blend_ok(eng, 128, 10 ** 5)                  # 4.2 MB per patch: over the 16 GB budget
    # bytes of the buffer: 64 x 65 x 128 channels at level 2, 32 x 33 x (128 + 256 injected) at level 3
    assert cfg.block_in_channels(128) == 128 and cfg.block_in_channels(64) == 384
    os.environ['NBE_BLEND_WAVEFRONT_GRAPHS'] = '1'
    try:
        assert not stylizer._phased_blend_ok(eng, 64, 2209)
    finally:
        del os.environ['NBE_BLEND_WAVEFRONT_GRAPHS']


def test_row_shards_match_shard_bounds_on_the_dense_grid():
    ys, xs = np.meshgrid(np.arange(47) * 88, np.arange(47) * 88, indexing='ij')
    yx = np.stack([ys.ravel(), xs.ravel()], axis=1).astype(np.int32)
    for world in (1, 2, 4, 8, 64):
        rows = stylizer.row_shards(47, world)
        assert [(a * 47, b * 47) if b > a else (len(yx), len(yx)) for a, b in rows] == \
            [(s, e) if e > s else (len(yx), len(yx)) for s, e in stylizer.shard_bounds(yx, world)]


def test_blending_wavefronts_respect_raster_dependencies():
    """Every raster-earlier patch whose window overlaps a patch lies in an earlier wavefront, and patches that share a
    wavefront do not overlap (so a wavefront can run as one batch)."""
    import numpy as np
    from brushstroke_engine_b200.stylizer import blending_wavefronts
    ys, xs = np.meshgrid(np.arange(6) * 108, np.arange(9) * 108, indexing='ij')
    yx = np.stack([ys.ravel(), xs.ravel()], axis=1)
    keep = np.ones(len(yx), dtype=bool); keep[[7, 20, 21, 33]] = False               # 'auto' stitching drops empty crops
    yx = yx[keep]
    waves = blending_wavefronts(yx, 128)
    assert sorted(np.concatenate(waves).tolist()) == list(range(len(yx)))
    wave_of = np.empty(len(yx), dtype=int)
    for w, idx in enumerate(waves):
        wave_of[idx] = w
    overlap = lambda a, b: abs(yx[a, 0] - yx[b, 0]) < 128 and abs(yx[a, 1] - yx[b, 1]) < 128
    for a in range(len(yx)):
        for b in range(a):
            if overlap(a, b):
                assert wave_of[b] < wave_of[a]
    assert len(waves) <= 2 * 6 + 9
    import pytest
    with pytest.raises(RuntimeError):
        blending_wavefronts(np.array([[0, 0], [0, 60]]), 128)                        # stride <= half a patch


def test_shard_bounds_equals_shard_crops():
    import numpy as np
    from brushstroke_engine_b200.stylizer import shard_bounds, shard_crops
    ys, xs = np.meshgrid(np.arange(11) * 108, np.arange(7) * 108, indexing='ij')
    yx = np.stack([ys.ravel(), xs.ravel()], axis=1)
    yx = yx[np.r_[0:20, 23:len(yx)]]                                      # a ragged row
    crops = [(int(y), int(x), 128, 128) for y, x in yx]
    for world in (1, 2, 3, 8, 16):
        assert shard_bounds(yx, world) == [shard_crops(crops, world, r) for r in range(world)]
    assert shard_bounds(np.zeros((0, 2), dtype=np.int32), 4) == [(0, 0)] * 4
