"""GPU parity of the engine composite and the batched patch scheduler against the reference-generated canvases
(tests/golden/engine.npz) and the CPU oracle: bit-exact crop list / tile placement, uint8 pixels within
1 LSB (FP32 mode; truncation can flip at a boundary) or 3 LSB (BF16 mode, 255 * 1e-2)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, t
from oracle import neube_oracle as O
from brushstroke_engine_b200 import params as P, synthetic

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(scope='module')
def engines(bundles):
    from brushstroke_engine_b200.engine import TriadPaintEngine
    cfg, ecfg, gp, ep = bundles
    return {m: TriadPaintEngine(gp, ep, DEV, mode=m) for m in ('fp32', 'bf16')}


def _opts(z, style_id=None):
    from brushstroke_engine_b200.engine import GanBrushOptions
    o = GanBrushOptions()
    o.set_style(z, style_id)
    return o


@pytest.mark.parametrize('mode,tol', [('fp32', 1e-4), ('bf16', 2e-2)])
@pytest.mark.parametrize('render_mode', ['clear', 'full'])
def test_render_stroke_torch_matches_oracle(engines, bundles, mode, tol, render_mode):
    cfg, ecfg, gp, ep = bundles
    eng = engines[mode]
    eng.set_render_mode(render_mode)
    geom = torch.from_numpy(np.concatenate([synthetic.synthetic_patch(128, seed=s) for s in (3, 4, 5)]))
    z = torch.cat([P.style_z_from_seed(s) for s in (594, 7, 11)])
    pos = torch.tensor([[0, 88], [264, 1144], [88, 88]])
    o = _opts(z.to(DEV))
    o.position = pos.to(DEV)
    o.set_color(1, np.array([255, 0, 128], dtype=np.uint8))          # user colour override (brush.py:514-527)
    rgba, raw, dbg = eng._render_stroke_torch(geom.to(DEV), None, o)
    gf = O.geometry_encode(ep, ecfg, geom)
    _, d = O.generator_forward(gp, cfg, z, gf, positions=pos)
    ref = O.triad_composite(d['uvs'], d['colors'], render_mode, color1=torch.tensor([1.0, 0.0, 128 / 255]))
    assert rgba.shape == (3, 4, 128, 128) and dbg is None
    assert float((rgba.cpu() - ref).abs().max()) < tol
    assert set(raw.keys()) >= {'uvs', 'colors', 'ws'}
    eng.set_render_mode('clear')
    with pytest.raises(RuntimeError):
        eng.set_render_mode('bogus')


@pytest.mark.parametrize('mode,lsb', [('fp32', 1), ('bf16', 3)])
@pytest.mark.parametrize('level,render_mode', [(0, 'clear'), (0, 'full'), (2, 'clear')])
def test_stylize_matches_reference_canvas(engines, mode, lsb, level, render_mode):
    from brushstroke_engine_b200 import stylizer
    g = load_golden('engine')
    eng = engines[mode]
    eng.set_render_mode(render_mode)
    out, job = stylizer.stylize(eng, g['guidance'], _opts(P.style_z_from_seed(594), '594'), crop_margin=10,
                                feature_blending_level=level, batch_size=5, return_job=True)
    eng.set_render_mode('clear')
    assert np.array_equal(np.array(job.crops, dtype=np.int32), g['crops'])            # bit-exact patch indexing
    assert np.array_equal(job.tiles_yx, g['metas'])                                   # bit-exact tile placement offsets
    ref = g[f'canvas_l{level}_{render_mode}']
    assert out.shape == ref.shape and out.dtype == np.uint8
    diff = np.abs(out.astype(np.int32) - ref.astype(np.int32))
    assert diff.max() <= lsb, diff.max()
    if mode == 'fp32':
        assert (diff > 0).mean() < 2e-3


def test_batching_is_invisible(engines):
    """Any batch size gives the same canvas bit for bit (ownership map, not launch order, resolves overlaps)."""
    from brushstroke_engine_b200 import stylizer
    g = load_golden('engine')
    eng = engines['bf16']
    o = _opts(P.style_z_from_seed(594), '594')
    a = stylizer.stylize(eng, g['guidance'], o, batch_size=16)
    b = stylizer.stylize(eng, g['guidance'], o, batch_size=3)
    c = stylizer.stylize(eng, g['guidance'], o, batch_size=1, on_white=True)
    assert np.array_equal(a, b)
    assert np.array_equal(stylizer.composite_on_white(a), c)


def test_owner_map_equals_raster_loop(engines):
    from brushstroke_engine_b200 import stylizer
    guidance = synthetic.synthetic_guidance(500, 333, num_lines=6, seed=2, radii=(3, 9))
    for mode in ('all', 'full'):
        job = stylizer.CanvasJob(engines['bf16'], guidance, 10, mode)
        owner = job.owner_map().cpu().numpy()
        ref = np.full((job.canvas_h, job.canvas_w), -1, dtype=np.int32)
        for i, (y, x) in enumerate(job.tiles_yx):
            ref[y:y + job.tile, x:x + job.tile] = i
        assert np.array_equal(owner, ref), mode
        crops_o, _ = O.generate_stitching_crops(O.pad_geo(guidance, 10), 128, mode, 20)
        assert crops_o == job.crops


def test_gather_geom_matches_prepare_geom_input(engines):
    from brushstroke_engine_b200 import stylizer
    guidance = synthetic.synthetic_guidance(300, 260, num_lines=10, seed=5, radii=(1, 3, 9))
    job = stylizer.CanvasJob(engines['bf16'], guidance, 10, 'all')
    got = job.gather(0, len(job.crops)).cpu()
    for i, (y, x, _, _) in enumerate(job.crops):
        ref = O.prepare_geom_input(255 - job.geom[y:y + 128, x:x + 128, :])
        assert torch.equal(got[i:i + 1], ref)
    e = engines['bf16']
    assert torch.equal(e.prepare_geom_input(255 - job.geom[0:128, 0:128, :]).cpu(), O.prepare_geom_input(255 - job.geom[0:128, 0:128, :]))


def test_uvs_mapping(engines, bundles):
    cfg, ecfg, gp, ep = bundles
    g = load_golden('engine')
    eng = engines['fp32']
    o = _opts(P.style_z_from_seed(594), '594')
    sf = eng.uvs_mapper.get_sfactor(o)                     # same synthetic mapper geometry as make_golden.py
    assert abs(float(sf) - float(g['sfactor'])) < 1e-4 * float(g['sfactor'])
    o.enable_uvs_mapping = True
    geom = torch.from_numpy(synthetic.synthetic_patch(128, seed=3))
    rgba, raw, _ = eng._render_stroke_torch(geom.to(DEV), None, o)
    ref = O.triad_composite(raw['uvs'].cpu(), raw['colors'].cpu(), 'clear', sfactor=torch.tensor(float(sf)))
    assert float((rgba.cpu() - ref).abs().max()) < 1e-5


def test_render_patches_host_equals_device_path(engines):
    eng = engines['bf16']
    guidance = synthetic.synthetic_guidance(300, 260, num_lines=10, seed=5, radii=(1, 3, 9))
    from brushstroke_engine_b200 import stylizer
    job = stylizer.CanvasJob(eng, guidance, 10, 'all')
    n = 6
    patches = torch.from_numpy(np.stack([job.geom[y:y + 128, x:x + 128, 0] for (y, x, _, _) in job.crops[:n]]))
    z = torch.cat([P.style_z_from_seed(i) for i in range(n)])
    pos = torch.from_numpy(job.crops_yx[:n].astype(np.int64))
    host = eng.render_patches_host(patches.pin_memory(), z.pin_memory(), pos.pin_memory(), crop_margin=10)
    o = _opts(z.to(DEV))
    o.position = pos.to(DEV)
    tiles, _ = eng.render_tiles(job.gather(0, n), o, crop_margin=10)
    assert host.shape == (n, 108, 108, 4) and torch.equal(host, tiles.cpu())
    # pipelined form: two pinned buffers, the copy of one call may still be in flight while the next call computes
    bufs = [torch.empty((n, 108, 108, 4), dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    evs = [eng.render_patches_host(patches.pin_memory(), z.pin_memory(), pos.pin_memory(), crop_margin=10, out=b, wait=False)[1] for b in bufs]
    for ev, b in zip(evs, bufs):
        ev.synchronize()
        assert torch.equal(b, host)


@pytest.mark.parametrize('mode', ['fp32', 'bf16'])
@pytest.mark.parametrize('level', [1, 2, 3])
def test_blending_wavefronts_equal_raster_loop(engines, mode, level):
    """The wavefront-batched feature-blending scheduler reproduces the sequential raster loop bit for bit
    (brush.py:190-242): same kernels on the same data, only the launch grouping differs."""
    from brushstroke_engine_b200 import stylizer
    eng = engines[mode]
    guidance = synthetic.synthetic_guidance(450, 560, num_lines=24, seed=3)          # 5 x 6 crops
    opts = _opts(P.style_z_from_seed(17), '17')
    job = stylizer.CanvasJob(eng, guidance, 10, 'all')
    assert len(job.crops) >= 20
    with torch.no_grad():
        seq = stylizer._stylize_blended(eng, job, opts, level, None)
        wav = stylizer._stylize_blended_wavefront(eng, job, opts, level, None)
    assert torch.equal(seq, wav)
    assert int((wav[..., 3] > 0).sum()) > 1000                                       # something was painted


def test_phased_blending_equals_wavefronts_and_raster_loop(engines):
    """Feature blending split at the blend point (every patch's layers before / after it at full batch size, only the blend kernel
    in wavefront order; stylizer._stylize_blended_phased) gives the bytes of the wavefront schedule and of the raster loop -- with a
    distinct style per patch and chunks smaller than the canvas, so that chunk and wavefront boundaries do not coincide."""
    from brushstroke_engine_b200 import stylizer
    eng = engines['bf16']
    guidance = synthetic.synthetic_guidance(450, 560, num_lines=24, seed=5)          # 5 x 6 crops
    opts = _opts(P.style_z_from_seed(17), '17')
    job = stylizer.CanvasJob(eng, guidance, 10, 'all')
    n = len(job.crops)
    zpp = torch.cat([P.style_z_from_seed(100 + i % 7) for i in range(n)]).to(eng.device)
    # level 2 blends at 64^2, level 3 at 32^2 (the block whose output buffer also carries the encoder's 256 injected channels),
    # level 1 at the output resolution (only ToRGB and the composite follow the blend)
    assert all(stylizer._phased_blend_ok(eng, r, n) for r in (128, 64, 32))
    with torch.no_grad():
        for level, z in ((2, None), (2, zpp), (3, zpp), (1, zpp)):
            seq = stylizer._stylize_blended_flat(eng, job, opts, level, z, sequential=True)
            wav = stylizer._stylize_blended_flat(eng, job, opts, level, z, sequential=False)
            for bs in (7, 256):
                ph = stylizer._stylize_blended_phased(eng, job, opts, level, z, batch_size=bs)
                assert torch.equal(ph, wav) and torch.equal(ph, seq), (level, z is None, bs)
    assert int((ph[..., 3] > 0).sum()) > 1000
    # a sparse crop list (crops without strokes are dropped: irregular wavefronts, fewer patches than grid cells)
    sparse = synthetic.synthetic_guidance(700, 610, num_lines=5, seed=11, radii=(3, 9))
    job2 = stylizer.CanvasJob(eng, sparse, 10, 'full')
    assert 2 <= len(job2.crops) < len(stylizer.CanvasJob(eng, sparse, 10, 'all').crops)
    with torch.no_grad():
        seq = stylizer._stylize_blended_flat(eng, job2, opts, 2, None, sequential=True)
        ph = stylizer._stylize_blended_phased(eng, job2, opts, 2, None, batch_size=4)
    assert torch.equal(ph, seq)


def test_interactive_graph_session_equals_render_stroke(engines):
    """The CUDA-graph replay of the batch-1 forward gives the same bytes as the eager call, for changing stroke patches
    and canvas positions (the shifted noise depends on the position and must be recomputed inside the graph)."""
    eng = engines['bf16']
    rng = np.random.RandomState(3)
    opts = _opts(P.style_z_from_seed(21), '21')
    sess = eng.interactive_session(opts, crop_margin=0)
    for k in range(4):
        geom = synthetic.synthetic_patch(128, seed=10 + k)[0, 0]
        patch = np.ascontiguousarray(((1.0 - geom) * 255).astype(np.uint8)[:, :, None])
        pos = (int(rng.randint(0, 3000)), int(rng.randint(0, 3000)))
        o = _opts(P.style_z_from_seed(21), '21')
        o.position = torch.tensor([[pos[0], pos[1]]], dtype=torch.int64)
        ref, _ = eng.render_stroke(patch, None, o)
        got = sess.render_stroke(patch, pos)
        assert got.shape == ref.shape == (128, 128, 4) and np.array_equal(got, ref), k
    # a new brush re-captures the graph
    opts2 = _opts(P.style_z_from_seed(99), '99')
    sess.set_brush(opts2)
    o = _opts(P.style_z_from_seed(99), '99')
    o.position = torch.tensor([[5, 7]], dtype=torch.int64)
    ref, _ = eng.render_stroke(patch, None, o)
    assert np.array_equal(sess.render_stroke(patch, (5, 7)), ref)


@pytest.mark.parametrize('B', [3, 40])
def test_batch_graph_session_equals_eager_render_tiles(engines, B):
    """The CUDA graph of the batch step (whole, and split before the last layer for the benchmark's kernel probe) returns the
    same bytes as the eager launch sequence, for changing inputs; so does render_patches_host, which feeds it from host buffers."""
    from brushstroke_engine_b200.engine import BatchSession
    eng = engines['bf16']
    rng = np.random.RandomState(B)
    whole = BatchSession(eng, B, 10)
    split = BatchSession(eng, B, 10, split_last_layer=True)
    assert whole.kernels_per_replay > split.kernels_per_replay > 20
    for rep in range(3):
        geom = torch.from_numpy(np.concatenate([synthetic.synthetic_patch(128, seed=100 * rep + i, radius=3 + (i % 4)) for i in range(B)])).to(eng.device)
        z = torch.cat([P.style_z_from_seed(1000 * rep + i) for i in range(B)]).to(eng.device)
        pos = torch.from_numpy(rng.randint(0, 4000, size=(B, 2))).to(eng.device)
        o = _opts(z)
        o.position = pos
        ref, _ = eng.render_tiles(geom, o, crop_margin=10)
        assert torch.equal(whole.run(geom, z, pos), ref), rep
        eng.G.probe = {'b128.conv1': []}
        got = split.run(geom, z, pos)
        probe, eng.G.probe = eng.G.probe['b128.conv1'], None
        assert torch.equal(got, ref) and len(probe) == 1
        torch.cuda.synchronize()
        assert probe[0][0].elapsed_time(probe[0][1]) > 0
        assert torch.equal(eng.render_tiles_graph(geom, z, pos, 10), ref)
        # host-fed: uint8 guidance patches (0 = stroke) -> pinned in, pinned out
        patches = (geom[:, 0] * 255).round().to(torch.uint8).cpu()
        g2 = (1 - (255 - patches.to(eng.device).float()) / 255.0)[:, None].contiguous()
        ref2, _ = eng.render_tiles(g2, o, crop_margin=10)
        out = eng.render_patches_host(patches.pin_memory(), z.cpu().pin_memory(), pos.cpu().pin_memory(), crop_margin=10)
        assert np.array_equal(out.numpy(), ref2.cpu().numpy()), rep


def test_interactive_graph_session_survives_other_batch_sizes(engines):
    """A graph session keeps replaying correctly while the same engine serves other batch sizes in between (the generator /
    encoder workspace caches are LRU: more distinct sizes than they hold evict the batch-1 buffers the graph points into,
    which the session must keep alive itself)."""
    eng = engines['bf16']
    opts = _opts(P.style_z_from_seed(33), '33')
    sess = eng.interactive_session(opts, crop_margin=0)
    geom = synthetic.synthetic_patch(128, seed=4)[0, 0]
    patch = np.ascontiguousarray(((1.0 - geom) * 255).astype(np.uint8)[:, :, None])
    o = _opts(P.style_z_from_seed(33), '33')
    o.position = torch.tensor([[40, 50]], dtype=torch.int64)
    ref, _ = eng.render_stroke(patch, None, o)
    assert np.array_equal(sess.render_stroke(patch, (40, 50)), ref)
    n_sizes = eng.G.max_cached_batch_sizes + 2
    for B in range(2, 2 + n_sizes):                               # evicts batch 1 from both caches
        g = torch.from_numpy(np.concatenate([synthetic.synthetic_patch(128, seed=50 + i) for i in range(B)])).to(eng.device)
        ob = _opts(torch.cat([P.style_z_from_seed(100 + i) for i in range(B)]))
        eng.render_tiles(g, ob, crop_margin=10)
    assert 1 not in eng.G._flat_ws
    torch.cuda.synchronize()
    junk = [torch.full((1 << 22,), 7.0, device=eng.device) for _ in range(8)]      # would land in freed batch-1 buffers
    assert np.array_equal(sess.render_stroke(patch, (40, 50)), ref)
    del junk
    ref2, _ = eng.render_stroke(patch, None, o)                   # a fresh batch-1 workspace next to the session's
    assert np.array_equal(ref2, ref) and np.array_equal(sess.render_stroke(patch, (40, 50)), ref)


@pytest.mark.parametrize('mode,tol', [('fp32', 1e-4), ('bf16', 2e-2)])
def test_wplus_library_style_with_noise_buffers_matches_oracle(engines, bundles, mode, tol, tmp_path):
    """A projected brush (w+ code and its own per-layer noise maps, forger/ui/library.py:146-202) set through
    ``WBrushLibrary`` drives the engine exactly as it drives the oracle: the noise buffers replace the shifted
    ``noise_const`` (networks.py:365-370), the mapping network is bypassed."""
    import pickle
    from brushstroke_engine_b200.library import WBrushLibrary
    from brushstroke_engine_b200.engine import GanBrushOptions
    cfg, ecfg, gp, ep = bundles
    eng = engines[mode]
    g = torch.Generator().manual_seed(11)
    def style():
        noise = {}
        for res in cfg.block_resolutions:
            for conv in (('conv0', 'conv1') if res > 4 else ('conv1',)):
                noise[f'b{res}.{conv}.noise_const'] = torch.randn(res, res, generator=g)
        return {'w': torch.randn(1, cfg.num_ws, cfg.w_dim, generator=g), 'noise': noise}
    path = str(tmp_path / 'brushes.pkl')
    with open(path, 'wb') as f:
        pickle.dump({'p': style(), 'q': style()}, f)
    lib = WBrushLibrary.from_file(path)
    o = GanBrushOptions()
    lib.set_interpolated_style('p', 'q', 0.35, o)
    ws, nb = o.style_ws.clone(), {k: v.clone() for k, v in o.custom_args['noise_buffers'].items()}
    geom = torch.from_numpy(np.concatenate([synthetic.synthetic_patch(128, seed=s) for s in (8, 9)]))
    pos = torch.tensor([[0, 88], [264, 1144]])
    o.position = pos.to(DEV)
    rgba, raw, _ = eng._render_stroke_torch(geom.to(DEV), None, o)
    gf = O.geometry_encode(ep, ecfg, geom)
    _, d = O.generator_forward(gp, cfg, None, gf, positions=pos, ws=ws.expand(2, -1, -1), noise_buffers=nb)
    ref = O.triad_composite(d['uvs'], d['colors'], 'clear')
    assert float((rgba.cpu() - ref).abs().max()) < tol
    # the noise maps matter: the same w+ without them gives a different image
    o2 = GanBrushOptions()
    o2.set_style_w(ws, style_id='x')
    o2.position = pos.to(DEV)
    rgba2, _, _ = eng._render_stroke_torch(geom.to(DEV), None, o2)
    assert float((rgba2 - rgba).abs().max()) > 1e-3


def test_uvs_mapper_batched_sfactors_icons_and_colors(engines):
    eng = engines['fp32']
    from brushstroke_engine_b200.engine import StyleUVSMapper
    single, batched = StyleUVSMapper(eng), StyleUVSMapper(eng)
    opts = [_opts(P.style_z_from_seed(s), str(s)) for s in (3, 594, 12)]
    ref = [single.get_sfactor(o) for o in opts]
    got = batched.get_sfactors(opts)
    for a, b in zip(ref, got):
        assert abs(float(a) - float(b)) <= 1e-5 * abs(float(a))
    assert all(str(s) in batched.sfactors for s in (3, 594, 12))
    assert batched.get_sfactors(opts[:1])[0] is batched.sfactors['3']                # served from the cache
    spec = single.get_colors(opts[1])
    assert spec.count('rgb(') == 3 and spec.count(':') == 2
    icon = single.get_brush_icon(opts[1])
    assert icon.shape == (128, 128, 3) and icon.dtype == np.uint8 and icon.std() > 0


@pytest.mark.parametrize('mode,tol', [('fp32', 1e-4), ('bf16', 2e-2)])
def test_canvas_colour_format_engine(bundles, mode, tol):
    """'canvas' colour format (networks.py:433-481) + CanvasPaintEngine's four render modes (brush.py:870-935) against the
    reference-generated fixture (FP32) / the oracle (BF16), and the batched uint8 tiles."""
    from brushstroke_engine_b200.engine import CanvasPaintEngine, GanBrushOptions
    _, ecfg, _, ep = bundles
    g = load_golden('canvas')
    cfg = P.GeneratorConfig(color_format='canvas')
    gp = P.init_generator_params(cfg, seed=3, perturb=0.1)
    assert P.bundle_digest(gp) == bytes(g['gen_digest']).decode()
    eng = CanvasPaintEngine(gp, ep, DEV, mode=mode)
    z, geom = t(g['z']), t(g['geom'])
    o = GanBrushOptions()
    o.set_style(z.to(DEV))
    o.set_color(1, np.array([255, 0, 128], dtype=np.uint8))
    for rm in ('clear', 'stroke', 'canvas', 'full'):
        eng.set_render_mode(rm)
        rgba, raw, _ = eng._render_stroke_torch(geom.to(DEV), None, o)
        assert float((rgba.cpu()[:, :, ::2, ::2] - t(g[f'rgba_{rm}_sub'])).abs().max()) < tol, rm
        tiles, _ = eng.render_tiles(geom.to(DEV), o, crop_margin=10)
        ref_u8 = O.to_uint8_tile(rgba.cpu(), 10)
        assert tiles.shape == (2, 108, 108, 4) and np.abs(tiles.cpu().numpy().astype(int) - ref_u8.astype(int)).max() <= 1
    assert float((raw['canvas'].cpu()[:, :, ::2, ::2] - t(g['canvas32_sub'])).abs().max()) < tol * 4       # un-normalised logits
    assert float((raw['alpha'].cpu()[:, :, ::2, ::2] - t(g['alpha32_sub'])).abs().max()) < tol
    assert set(raw.keys()) >= {'uvs', 'colors', 'canvas', 'alpha_fg', 'alpha', 'ws'}
    with pytest.raises(RuntimeError):
        eng.set_render_mode('bogus')
    with pytest.raises(RuntimeError):
        CanvasPaintEngine(gp, ep, DEV, mode=mode, gen_cfg=P.GeneratorConfig())


def test_sparse_stitching_mode_keeps_the_oracles_crops(engines):
    """Stitching modes other than 'all' keep the crops with more than 10 stroke pixels (style_transfer.py:43-47); the
    window counts come from nbe_count_stroke_pixels."""
    from brushstroke_engine_b200 import stylizer
    eng = engines['bf16']
    guidance = synthetic.synthetic_guidance(700, 900, num_lines=3, seed=2, radii=(2, 3))      # sparse: many empty crops
    job = stylizer.CanvasJob(eng, guidance, 10, 'stroke')
    ref_crops, _ = O.generate_stitching_crops(O.pad_geo(guidance, 10), 128, 'stroke', 20)
    assert job.crops == ref_crops
    full = stylizer.CanvasJob(eng, guidance, 10, 'all')
    assert 0 < len(job.crops) < len(full.crops)
    out = stylizer.stylize(eng, guidance, _opts(P.style_z_from_seed(5), '5'), crop_margin=10, stitching_mode='stroke')
    assert out.shape == (700, 900, 4) and int((out[..., 3] > 0).sum()) > 500


def test_fp32_reflect_pad_and_bilinear_kernel_matches_torch():
    import torch.nn.functional as F
    from brushstroke_engine_b200 import _lib
    g = torch.Generator().manual_seed(0)
    for (n, c, h, w, pad, up) in [(2, 3, 9, 7, 1, 0), (1, 5, 16, 16, 3, 0), (3, 4, 8, 8, 1, 1), (2, 2, 5, 11, 1, 1)]:
        x = torch.randn(n, c, h, w, generator=g).to(DEV)
        S = 2 if up else 1
        y = torch.empty((n, c, S * h + 2 * pad, S * w + 2 * pad), device=DEV)
        _lib.call('nbe_reflect_pad_nchw_f32', _lib.ptr(x), _lib.ptr(y), n * c, h, w, pad, up, _lib.stream())
        ref = F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=True) if up else x
        ref = F.pad(ref, (pad, pad, pad, pad), mode='reflect')
        assert float((y - ref).abs().max()) < (2e-6 if up else 0.0) + 1e-12, (n, c, h, w, pad, up)
