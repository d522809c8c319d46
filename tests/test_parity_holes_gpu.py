"""GPU: the branches round 1 left without a test -- truncation_psi / truncation_cutoff, encoder preproc_type, the
``attach_fast_path`` / ``engine_from_reference`` hooks (through a stand-in reference engine), the band-sharded canvas against
the single-GPU canvas, and (with >= 2 devices) the NCCL exchange under torchrun."""
import dataclasses
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import REPO, load_golden, t
from oracle import neube_oracle as O
from brushstroke_engine_b200 import params as P, synthetic

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def md(a, b):
    return float((torch.as_tensor(a).detach().cpu().double() - torch.as_tensor(b).detach().cpu().double()).abs().max())


@pytest.mark.parametrize('psi,cutoff', [(0.7, None), (0.5, 8), (1.3, 3)])
def test_truncation_psi_and_cutoff(bundles, psi, cutoff):
    """MappingNetwork truncation (networks.py:283-289): ws = w_avg.lerp(ws, psi), optionally on the first `cutoff` layers
    only -- with a non-zero w_avg (the constructor's zeros would hide a wrong branch)."""
    from brushstroke_engine_b200.generator import Generator
    cfg, ecfg, gp, ep = bundles
    gp = dict(gp)
    gp['mapping.w_avg'] = torch.randn(cfg.w_dim, generator=torch.Generator().manual_seed(5)) * 0.5
    G = Generator(gp, cfg, DEV, mode='fp32')
    z = torch.cat([P.style_z_from_seed(s) for s in (594, 7, 21)])
    ws = G.mapping(z.to(DEV), None, truncation_psi=psi, truncation_cutoff=cutoff)
    ref = O.mapping_network(gp, z, cfg.num_ws, cfg.mapping_layers, cfg.mapping_lr_multiplier, truncation_psi=psi, truncation_cutoff=cutoff)
    assert ws.shape == ref.shape == (3, cfg.num_ws, cfg.w_dim)
    assert md(ws, ref) < 1e-5
    # and through the full call: psi changes the image, psi = 1 does not
    g = load_golden('generator')
    gf = [x.to(DEV) for x in O.geometry_encode(ep, ecfg, t(g['geom']))]
    img1 = G(z[:2].to(DEV), None, gf, noise_mode='const', truncation_psi=1)
    img2 = G(z[:2].to(DEV), None, gf, noise_mode='const', truncation_psi=psi, truncation_cutoff=cutoff)
    gf_c = O.geometry_encode(ep, ecfg, t(g['geom']))
    ref2, _ = O.generator_forward(gp, cfg, z[:2], gf_c, ws=ref[:2])
    assert md(img2, ref2) < 1e-4 and md(img1, img2) > 1e-3


@pytest.mark.parametrize('preproc', ['inverse', '-11inverse', 'none'])
def test_encoder_preproc_types(bundles, preproc):
    """BaseGeoEncoder preprocessing (base.py:32-58) in both precision modes against the oracle."""
    from brushstroke_engine_b200.geo_encoder import GeometryEncoder
    cfg, ecfg, gp, ep = bundles
    e2 = dataclasses.replace(ecfg, preproc_type=preproc)
    geom = torch.from_numpy(np.concatenate([synthetic.synthetic_patch(128, seed=3), synthetic.synthetic_patch(128, seed=9, radius=5)]))
    ref = O.geometry_encode(ep, e2, geom)
    if preproc != 'none':
        assert md(ref[0], O.geometry_encode(ep, ecfg, geom)[0]) > 1e-3          # the flag matters for these weights
    got = GeometryEncoder(ep, e2, DEV, mode='fp32').encode(geom.to(DEV))
    for a, b in zip(got, ref):
        assert md(a, b) < 5e-5, preproc
    got = GeometryEncoder(ep, e2, DEV, mode='bf16').encode(geom.to(DEV))
    for a, b in zip(got, ref):
        assert md(a, b) < 0.03 * float(b.abs().max()), preproc


@pytest.mark.parametrize('mode,tol', [('fp32', 1e-4), ('bf16', 2e-2)])
def test_attach_fast_path_on_a_reference_shaped_engine(bundles, mode, tol):
    """``attach_fast_path(engine)`` reads G / encoder through ``state_dict()`` and the module
    attributes, builds the B200 engine and replaces ``_render_stroke_torch`` -- executed here on a stand-in engine whose
    modules are laid out like the reference's (tests/standin.py), and compared with the oracle composite."""
    from standin import standin_engine
    from brushstroke_engine_b200 import install
    from brushstroke_engine_b200.engine import GanBrushOptions, TriadPaintEngine
    cfg, ecfg, gp, ep = bundles
    ref_engine = standin_engine(cfg, ecfg, gp, ep, DEV, render_mode='full')
    fast = install.attach_fast_path(ref_engine, mode=mode)
    assert isinstance(fast, TriadPaintEngine) and ref_engine._nbe_fast is fast and fast.G.cfg == cfg and fast.encoder.cfg == ecfg
    geom = torch.from_numpy(np.concatenate([synthetic.synthetic_patch(128, seed=3), synthetic.synthetic_patch(128, seed=4, radius=3)]))
    z = torch.cat([P.style_z_from_seed(594), P.style_z_from_seed(7)])
    pos = torch.tensor([[88, 176], [1144, 264]])
    gf = O.geometry_encode(ep, ecfg, geom)
    _, dbg = O.generator_forward(gp, cfg, z, gf, positions=pos)
    for render_mode in ('full', 'clear'):
        ref_engine.render_mode = render_mode                         # the hook follows the reference engine's mode
        opts = GanBrushOptions()
        opts.set_style(z.to(DEV))
        opts.position = pos.to(DEV)
        rgba, raw, dbg_img = ref_engine._render_stroke_torch(geom.to(DEV), None, opts)
        assert md(rgba, O.triad_composite(dbg['uvs'], dbg['colors'], render_mode)) < tol, render_mode
        assert set(('uvs', 'colors', 'ws')) <= set(raw.keys()) and dbg_img is None
    # UVS mapping: the hook takes the reference mapper's factor (it depends on the reference's bundled geometry images)
    ref_engine.uvs_mapper.get_sfactor = lambda o: torch.tensor(1.7)
    opts = GanBrushOptions()
    opts.set_style(z[:1].to(DEV), 'brush-1')
    opts.enable_uvs_mapping = True
    rgba, _, _ = ref_engine._render_stroke_torch(geom[:1].to(DEV), None, opts)
    _, dbg1 = O.generator_forward(gp, cfg, z[:1], [x[:1] for x in gf])
    assert md(rgba, O.triad_composite(dbg1['uvs'], dbg1['colors'], 'clear', sfactor=torch.tensor(1.7))) < tol
    eng2 = install.engine_from_reference(ref_engine, mode=mode)
    assert eng2.render_mode == ref_engine.render_mode


@pytest.fixture(scope='module')
def v2_bundle():
    ecfg2 = P.EncoderConfig(bn_after_activation=True, neg_slope=0.2)
    ep2 = P.init_encoder_params(ecfg2, seed=5, perturb_bn=0.1)
    g = load_golden('encoder_v2')
    assert bytes(g['enc_digest']).decode() == P.bundle_digest(ep2), 'encoder weights differ from the golden run'
    return ecfg2, ep2, g


def test_encoder_neg_slope_variant_fp32(v2_bundle):
    """The --neg_slope autoencoder (conv -> LeakyReLU -> BatchNorm, ScaleUpV2; simple_autoencoder.py:48-53,128-148) in FP32
    parity mode against the features of the reference's own factory-built model (tests/golden/encoder_v2.npz)."""
    from brushstroke_engine_b200.geo_encoder import GeometryEncoder
    ecfg2, ep2, g = v2_bundle
    enc = GeometryEncoder(ep2, ecfg2, DEV, mode='fp32')
    g0, g1 = enc.encode(t(g['geom']).to(DEV))
    assert g0.shape == (2, 16, 16, 16) and g1.shape == (2, 256, 32, 32)
    assert md(g0, g['g0']) < 5e-5
    assert md(g1[:, ::8], g['g1_sub']) < 5e-5
    # one resolution only, and another input size (the fold of BatchNorm into the next conv relies on reflect padding only)
    only0 = enc.encode(t(g['geom']).to(DEV), res=0)
    assert md(only0[0] if isinstance(only0, (list, tuple)) else only0, g['g0']) < 5e-5
    geom64 = t(g['geom'])[:, :, 32:96, 32:96].contiguous()
    ref = O.geometry_encode(ep2, ecfg2, geom64)
    got = enc.encode(geom64.to(DEV))
    for a, b in zip(got, ref):
        assert md(a, b) < 5e-5


def test_encoder_neg_slope_variant_bf16(v2_bundle):
    """The same on the tensor-core path: folded-forward BatchNorm, ScaleUpV2 as transposed conv on tcgen05 + crop / bias /
    LeakyReLU + BatchNorm; within bf16 rounding of the reference features, like the default layout's test."""
    from brushstroke_engine_b200.geo_encoder import GeometryEncoder
    ecfg2, ep2, g = v2_bundle
    ref = O.geometry_encode(ep2, ecfg2, t(g['geom']))
    enc = GeometryEncoder(ep2, ecfg2, DEV, mode='bf16')
    got = enc.encode(t(g['geom']).to(DEV))
    assert got[0].shape == (2, 16, 16, 16) and got[1].shape == (2, 256, 32, 32)
    for a, b in zip(got, ref):
        err = md(a, b)
        rms = float(((a.cpu().double() - b.double()) ** 2).mean().sqrt() / (b.double() ** 2).mean().sqrt())
        print(f'bf16 --neg_slope encoder: max-abs error {err:.4g} (max |ref| {float(b.abs().max()):.4g}), relative RMS error {rms:.3e}')
        assert err < 0.02 * float(b.abs().max()) and rms < 1e-2, (err, rms)
    geom5 = t(g['geom']).repeat(3, 1, 1, 1)[:5].to(DEV)
    h = enc.encode(geom5)
    assert md(h[0][:2], got[0]) == 0 and md(h[1][:2], got[1]) == 0


@pytest.mark.parametrize('mode,tol', [('fp32', 1e-4), ('bf16', 2e-2)])
def test_engine_with_neg_slope_encoder_through_the_hook(bundles, v2_bundle, mode, tol):
    """A reference-shaped engine whose encoder is the --neg_slope variant goes through ``attach_fast_path``: the hook reads the
    layout from the modules and the rendered stroke matches the oracle (generator fed with the oracle's v2 features)."""
    from standin import standin_engine, StandInAutoEncoder
    from brushstroke_engine_b200 import install
    from brushstroke_engine_b200.engine import GanBrushOptions
    cfg, ecfg, gp, ep = bundles
    ecfg2, ep2, g = v2_bundle
    ref_engine = standin_engine(cfg, ecfg, gp, ep, DEV)
    ref_engine.encoder = StandInAutoEncoder(ecfg2, ep2, scale_up_v2=True, neg_slope=0.2, bn_after_act=True)
    fast = install.attach_fast_path(ref_engine, mode=mode)
    assert fast.encoder.cfg == ecfg2
    geom = t(g['geom'])
    z = torch.cat([P.style_z_from_seed(594), P.style_z_from_seed(7)])
    pos = torch.tensor([[88, 176], [1144, 264]])
    gf = O.geometry_encode(ep2, ecfg2, geom)
    _, dbg = O.generator_forward(gp, cfg, z, gf, positions=pos)
    opts = GanBrushOptions()
    opts.set_style(z.to(DEV))
    opts.position = pos.to(DEV)
    rgba, raw, _ = ref_engine._render_stroke_torch(geom.to(DEV), None, opts)
    assert md(rgba, O.triad_composite(dbg['uvs'], dbg['colors'], 'clear')) < tol


def test_neg_slope_encoder_in_the_batch_graph(bundles, v2_bundle):
    """The batch step as a CUDA graph (what stylize / render_patches_host replay) with the --neg_slope encoder: same bytes as the
    eager launch sequence for changing inputs -- the variant's extra buffers (dense last activation, zero-gapped g0, transposed-conv
    result) live in the graph's pool / the encoder's workspace cache."""
    from brushstroke_engine_b200.engine import BatchSession, GanBrushOptions, TriadPaintEngine
    cfg, ecfg, gp, ep = bundles
    ecfg2, ep2, g = v2_bundle
    eng = TriadPaintEngine(gp, ep2, DEV, mode='bf16', enc_cfg=ecfg2)
    B = 5
    sess = BatchSession(eng, B, 10)
    rng = np.random.RandomState(7)
    for rep in range(3):
        geom = torch.from_numpy(np.concatenate([synthetic.synthetic_patch(128, seed=50 * rep + i, radius=3 + (i % 4)) for i in range(B)])).to(DEV)
        z = torch.cat([P.style_z_from_seed(100 * rep + i) for i in range(B)]).to(DEV)
        pos = torch.from_numpy(rng.randint(0, 4000, size=(B, 2))).to(DEV)
        o = GanBrushOptions()
        o.set_style(z)
        o.position = pos
        ref, _ = eng.render_tiles(geom, o, crop_margin=10)
        assert torch.equal(sess.run(geom, z, pos), ref), rep
    assert int(ref.max()) > 0


def test_band_jobs_reassemble_the_single_gpu_canvas(bundles):
    """The multi-GPU scheduler's per-rank work -- a row-window CanvasJob (partial guidance upload, global positions), its own
    tiles placed into the canvas rows it owns -- executed for every 'rank' in turn on ONE device: the concatenated bands must
    equal the canvas stylize() renders in one piece, bit for bit (the NCCL exchange itself only moves those bands)."""
    from brushstroke_engine_b200 import stylizer
    from brushstroke_engine_b200.engine import GanBrushOptions, TriadPaintEngine
    cfg, ecfg, gp, ep = bundles
    eng = TriadPaintEngine(gp, ep, DEV, mode='bf16')
    guidance = synthetic.synthetic_guidance(700, 610, num_lines=24, seed=3)
    opts = GanBrushOptions()
    opts.set_style(P.style_z_from_seed(594).to(DEV), '594')
    m = 10
    nrows, ncols, rwidth, ph, pw = stylizer.crop_grid(700 + m, 610 + m, 128, 2 * m)
    zpp = torch.cat([P.style_z_from_seed(i % 5) for i in range(nrows * ncols)]).to(DEV)
    with torch.no_grad():
        whole = stylizer.stylize(eng, guidance, opts, crop_margin=m, z_per_patch=zpp, to_host=False, distributed=False)
        for world in (2, 3, nrows + 2):
            canvas = torch.full((ph, pw, 4), 99, dtype=torch.uint8, device=DEV)
            for rank, (r0, r1) in enumerate(stylizer.row_shards(nrows, world)):
                job = stylizer.CanvasJob(eng, guidance if rank % 2 else torch.from_numpy(guidance).to(DEV), m, 'all', crop_rows=(r0, r1))
                tiles = stylizer._render_job_tiles(eng, job, opts, zpp, r0 * ncols, 64, m)
                lo, hi = job.band
                band = canvas[lo:hi]
                band.zero_()
                job.place_band(band, tiles)
            assert torch.equal(canvas[m:m + 700, m:m + 610], whole), world


def test_multi_gpu_canvas_equals_single_gpu_under_torchrun(tmp_path):
    """2 ranks over NCCL (skipped below 2 devices): stylize() sharded == stylize() on one GPU, dense and sparse grids, and a
    feature-blended canvas whose layers around the blend point are sharded (one style, then a style per patch)."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    script = tmp_path / 'mg.py'
    script.write_text('''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
from brushstroke_engine_b200 import params as P, stylizer, synthetic
from brushstroke_engine_b200.engine import GanBrushOptions, TriadPaintEngine
rank = int(os.environ['LOCAL_RANK']); torch.cuda.set_device(rank)
dev = torch.device('cuda', rank)
dist.init_process_group('nccl', device_id=dev)
cfg, ecfg = P.GeneratorConfig(), P.EncoderConfig()
eng = TriadPaintEngine(P.init_generator_params(cfg, 0, 0.1), P.init_encoder_params(ecfg, 1, 0.1), dev, mode='bf16')
guidance = synthetic.synthetic_guidance(900, 700, num_lines=30, seed=2)
opts = GanBrushOptions(); opts.set_style(P.style_z_from_seed(594).to(dev), '594')
ok = True
with torch.no_grad():
    for mode in ('all', 'full'):
        out = stylizer.stylize(eng, guidance, opts, crop_margin=10, stitching_mode=mode, to_host=False)
        if rank == 0:
            solo = stylizer.stylize(eng, guidance, opts, crop_margin=10, stitching_mode=mode, to_host=False, distributed=False)
            ok = ok and bool(torch.equal(out, solo))
        dist.barrier()
    # one feature-blended canvas (level 2) over both ranks: layers before / after the blend point sharded, the blend on rank 0
    zpp = None
    for rep in range(2):
        out = stylizer.stylize(eng, guidance, opts, crop_margin=10, feature_blending_level=2, z_per_patch=zpp, to_host=False, batch_size=16)
        if rank == 0:
            solo = stylizer.stylize(eng, guidance, opts, crop_margin=10, feature_blending_level=2, z_per_patch=zpp, to_host=False, distributed=False)
            ok = ok and bool(torch.equal(out, solo)) and int((out[..., 3] > 0).sum()) > 1000
        else:
            ok = ok and out is None
        dist.barrier()
        n = len(stylizer.CanvasJob(eng, guidance, 10, 'all').crops)
        zpp = torch.cat([P.style_z_from_seed(i %% 5) for i in range(n)]).to(dev)
if rank == 0:
    open(%r, 'w').write('OK' if ok else 'MISMATCH')
dist.destroy_process_group()
''' % (REPO, str(tmp_path / 'result.txt')))
    subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr', '127.0.0.1',
                    '--master-port', str(29700 + os.getpid() % 200), str(script)], check=True, timeout=600)
    assert (tmp_path / 'result.txt').read_text() == 'OK'


@pytest.mark.parametrize('mode', ['fp32', 'bf16'])
def test_styles_table_broadcast_view_and_row_ranges(bundles, mode):
    """The single styles launch (``nbe_styles_demod_f32``): a w repeated for every layer read through a stride-0 view gives the
    same bits as the reference's materialised ``repeat`` (networks.py:281), and every row-range entry (ToRGB colours / styles,
    ':head' / ':geo' halves of the concat layers) equals the slice of the whole layer's styles (networks.py:109-122, 455-462)."""
    from brushstroke_engine_b200.generator import Generator
    cfg, ecfg, gp, ep = bundles
    G = Generator(gp, cfg, DEV, mode=mode)
    z = torch.cat([P.style_z_from_seed(s) for s in range(11)]).to(DEV)
    ws_rep = G.mapping(z, None)
    ws_view = G.mapping(z, None, broadcast_view=True)
    assert ws_rep.is_contiguous() and ws_view.stride(1) == 0 and torch.equal(ws_rep, ws_view)
    assert G.mapping(z, None, truncation_psi=0.5, truncation_cutoff=4, broadcast_view=True).is_contiguous()   # written in place: materialised
    s_rep, d_rep, c_rep, r_rep = G._styles(ws_rep)
    s_view, d_view, c_view, r_view = G._styles(ws_view)
    assert set(s_rep) == set(s_view) and set(d_rep) == set(d_view)
    for k in s_rep:
        assert torch.equal(s_rep[k], s_view[k]), k
    for k in d_rep:
        assert torch.equal(d_rep[k], d_view[k]), k
    assert torch.equal(c_rep, c_view) and torch.equal(r_rep, r_view)
    # a genuinely per-layer w+ (every layer its own latent) through the same launch, against the oracle's affine layers
    wp = torch.randn(5, cfg.num_ws, cfg.w_dim, generator=torch.Generator().manual_seed(2)).to(DEV)
    s_wp, _, c_wp, r_wp = G._styles(wp)
    names = [k for k in s_wp if ':' not in k]
    for li, k in enumerate(names):
        ref = O.fully_connected(wp[:, li].cpu(), gp[f'synthesis.{k}.affine.weight'], gp[f'synthesis.{k}.affine.bias'])
        assert md(s_wp[k], ref) < 1e-4, k
    for k in s_wp:
        if k.endswith(':head'):
            base = s_wp[k.split(':')[0]]
            assert torch.equal(s_wp[k], base[:, :s_wp[k].shape[1]]), k
        if k.endswith(':geo'):
            base = s_wp[k.split(':')[0]]
            assert torch.equal(s_wp[k], base[:, base.shape[1] - s_wp[k].shape[1]:]), k
    assert c_wp.shape == (5, 3, 3) and r_wp.shape[0] == 5 and float(c_wp.abs().max()) <= 1.0


@pytest.mark.parametrize('B', [1, 11, 64])
def test_styles_launch_writes_the_modulated_constant_input(bundles, B):
    """``nbe_styles_demod_input_f32``: the b4 input ``const * styles(b4.conv1)`` (networks.py:642-643 + :68) written by the styles
    launch equals the torch expression bit for bit, the gap column stays untouched, and the flat path picks it up only for the
    very styles / buffer pair it was written for (full images equal with and without it)."""
    from brushstroke_engine_b200.generator import Generator
    cfg, ecfg, gp, ep = bundles
    G = Generator(gp, cfg, DEV, mode='bf16')
    ws = torch.randn(B, cfg.num_ws, cfg.w_dim, generator=torch.Generator().manual_seed(B)).to(DEV)
    in4 = torch.full((B, 4, 5, cfg.channels(4)), 7.0, dtype=torch.bfloat16, device=DEV)
    styles, _, _, _ = G._styles(ws, in4=in4)
    assert G._in4_filled is not None and G._in4_filled[0] is styles['b4.conv1'] and G._in4_filled[1] is in4
    ref = (G._const_nhwc.unsqueeze(0) * styles['b4.conv1'][:, None, None, :]).to(torch.bfloat16)
    assert torch.equal(in4[:, :, :4, :], ref)
    assert bool((in4[:, :, 4, :] == 7.0).all())
    plain, _, _, _ = G._styles(ws)
    assert G._in4_filled is None and torch.equal(plain['b4.conv1'], styles['b4.conv1'])
    if G.flat_supported:
        geom = [torch.randn(B, c, r, r, generator=torch.Generator().manual_seed(r)).to(DEV)
                for r, c in zip(cfg.geom_feature_resolutions, cfg.geom_feature_channels)]
        pos = torch.randint(0, 4096, (B, 2), generator=torch.Generator().manual_seed(3)).to(DEV)
        img_a = G.forward_pre_mapped(ws, geom, positions=pos, noise_mode='const')           # styles launch fills in4
        G._styles(ws, in4=G._workspace(B)['in4'])                                            # stale pair: other styles object
        real = G._styles
        G._styles = lambda w, in4=None: real(w)                                              # torch expression
        try:
            img_b = G.forward_pre_mapped(ws, geom, positions=pos, noise_mode='const')
        finally:
            G._styles = real
        img_a = img_a[0] if isinstance(img_a, tuple) else img_a
        img_b = img_b[0] if isinstance(img_b, tuple) else img_b
        assert torch.equal(img_a, img_b)
