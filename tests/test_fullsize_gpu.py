"""Parity at BASELINE.json's full sizes (batch 256 x 128x128 patches, 4096^2 canvas) through size-independent properties:
the small-size runs of the same kernels are checked element by element against the oracle / fp64 references elsewhere
(test_generator_gpu.py, test_conv_*_gpu.py, test_engine_gpu.py); here the FULL-size runs are tied to them.

* patches are independent units: a patch rendered inside a batch of 256 equals, bit for bit, the same patch rendered in a
  batch of 8 (so the oracle-checked small batches speak for the large ones), and permuting the batch permutes the output;
* the operators are homogeneous: scaling the input by a power of two scales the output exactly (bias_act 'linear',
  upfirdn2d), at the microbenchmark's full tensor sizes;
* the canvas scheduler is deterministic and every canvas pixel is the pixel of the tile that owns it (last writer wins),
  checked on the full 4096^2 canvas against a host re-computation of the ownership rule.
"""
import numpy as np
import pytest
import torch

from brushstroke_engine_b200 import params as P, synthetic

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(scope='module')
def engine(bundles):
    from brushstroke_engine_b200.engine import TriadPaintEngine
    cfg, ecfg, gp, ep = bundles
    return TriadPaintEngine(gp, ep, DEV, mode='bf16')


def _opts(z, pos):
    from brushstroke_engine_b200.engine import GanBrushOptions
    o = GanBrushOptions()
    o.set_style(z)
    o.position = pos
    return o


def _workload(B, seed=0):
    g = torch.Generator().manual_seed(seed)
    guidance = synthetic.synthetic_guidance(2000, 2000, num_lines=96, seed=seed)
    ys = torch.randint(0, 2000 - 128, (B,), generator=g)
    xs = torch.randint(0, 2000 - 128, (B,), generator=g)
    geom = torch.stack([torch.from_numpy(guidance[y:y + 128, x:x + 128, -1].astype(np.float32) / 255.0) for y, x in zip(ys.tolist(), xs.tolist())])
    z = torch.randn(B, 64, generator=g, dtype=torch.float64)
    pos = torch.stack([ys, xs], dim=1).to(torch.int64)
    return geom[:, None].contiguous().to(DEV), z.to(DEV), pos.to(DEV)


def test_batch_256_equals_small_batches_bit_for_bit(engine):
    geom, z, pos = _workload(256)
    with torch.no_grad():
        full, _ = engine.render_tiles(geom, _opts(z, pos), crop_margin=10)
        assert full.shape == (256, 108, 108, 4) and full.dtype == torch.uint8
        for s in (0, 120, 248):
            small, _ = engine.render_tiles(geom[s:s + 8].contiguous(), _opts(z[s:s + 8], pos[s:s + 8]), crop_margin=10)
            assert torch.equal(small, full[s:s + 8]), s
        assert int((full[..., 3] > 0).sum()) > 100000                    # the strokes were painted


def test_batch_256_permutation_equivariance(engine):
    geom, z, pos = _workload(256, seed=1)
    perm = torch.randperm(256, generator=torch.Generator().manual_seed(5)).to(DEV)
    with torch.no_grad():
        a, _ = engine.render_tiles(geom, _opts(z, pos), crop_margin=10)
        b, _ = engine.render_tiles(geom[perm].contiguous(), _opts(z[perm], pos[perm]), crop_margin=10)
    assert torch.equal(b, a[perm])


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_operators_are_homogeneous_at_full_size(dtype):
    from brushstroke_engine_b200 import upfirdn2d as U
    from brushstroke_engine_b200.bias_act import bias_act
    g = torch.Generator(device=DEV).manual_seed(3)
    f4 = U.setup_filter([1, 3, 3, 1], device=DEV)
    x = torch.randn(256, 128, 65, 65, device=DEV, dtype=dtype, generator=g)          # the generator's [2R+1]^2 -> [2R]^2 call, R = 32
    y1 = U.upfirdn2d(x, f4, padding=[1, 1, 1, 1], gain=4)
    y2 = U.upfirdn2d(x * 4, f4, padding=[1, 1, 1, 1], gain=4)
    assert y1.shape == (256, 128, 64, 64) and torch.equal(y2, y1 * 4)                # power-of-two scaling is exact in every dtype
    u1, u2 = U.upsample2d(x[:, :, :32, :32].contiguous(), f4), U.upsample2d((x[:, :, :32, :32] * 0.5).contiguous(), f4)
    assert u1.shape == (256, 128, 64, 64) and torch.equal(u2, u1 * 0.5)
    b = torch.randn(128, device=DEV, dtype=dtype, generator=g)
    z1 = bias_act(y1, b, act='linear')
    z2 = bias_act(y1 * 2, b * 2, act='linear')
    assert torch.equal(z2, z1 * 2)
    # lrelu is positively homogeneous too: lrelu(2 (x + b)) * g = 2 lrelu(x + b) * g, clamp disabled
    l1 = bias_act(y1, b, act='lrelu', gain=2.0)
    l2 = bias_act(y1 * 2, b * 2, act='lrelu', gain=2.0)
    assert torch.equal(l2, l1 * 2)


def test_canvas_4096_is_deterministic_and_owner_consistent(engine):
    from brushstroke_engine_b200 import stylizer
    guidance = synthetic.synthetic_guidance(4096, 4096, num_lines=256, seed=0)
    d_guidance = torch.from_numpy(guidance).to(DEV)
    opts = _opts(torch.from_numpy(np.random.RandomState(594).randn(1, 64)).to(DEV), None)
    with torch.no_grad():
        out1, job = stylizer.stylize(engine, d_guidance, opts, crop_margin=10, batch_size=256, to_host=False, return_job=True)
        out2 = stylizer.stylize(engine, d_guidance, opts, crop_margin=10, batch_size=192, to_host=False)
    assert out1.shape == (4096, 4096, 4) and torch.equal(out1, out2)                 # batch size is invisible at full size too
    assert len(job.crops) == 47 * 47
    # ownership rule on the host: the owner of a pixel is the LAST crop in raster order whose tile covers it
    m, T = 10, 108
    tys = np.unique(job.tiles_yx[:, 0]); txs = np.unique(job.tiles_yx[:, 1])
    rng = np.random.RandomState(0)
    py = rng.randint(m, 4096 + m, size=4000); px = rng.randint(m, 4096 + m, size=4000)   # padded-canvas coordinates
    oy = np.searchsorted(tys, py, side='right') - 1; ox = np.searchsorted(txs, px, side='right') - 1
    ok = (py - tys[oy] < T) & (px - txs[ox] < T)
    owner = oy * len(txs) + ox
    # re-render the owning patches and compare the sampled pixels
    sel = np.unique(owner[ok])[:64]
    d_idx = torch.from_numpy(sel.astype(np.int64)).to(DEV)
    with torch.no_grad():
        o = _opts(opts.style_z, job.d_crops[d_idx].to(torch.int64))
        tiles, _ = engine.render_tiles(job.gather_indices(d_idx), o, crop_margin=10)
    tiles = tiles.cpu().numpy(); canvas = out1.cpu().numpy()
    checked = 0
    for k, n in enumerate(sel):
        for j in np.flatnonzero(ok & (owner == n)):
            ty, tx = job.tiles_yx[n]
            assert np.array_equal(canvas[py[j] - m, px[j] - m], tiles[k, py[j] - ty, px[j] - tx]), (n, j)
            checked += 1
    assert checked > 50
