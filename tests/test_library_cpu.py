"""Brush libraries (host-side plumbing): file formats, style / interpolation rules, and -- when the reference
is present (build container only) -- equality with forger/ui/library.py on the same inputs."""
import os
import pickle
import sys
import types

import numpy as np
import pytest
import torch

from brushstroke_engine_b200 import library as L
from brushstroke_engine_b200.engine import GanBrushOptions

REF = '/root/reference'


def _wlib(tmp_path):
    g = torch.Generator().manual_seed(0)
    styles = {
        'a': {'w': torch.randn(1, 14, 64, generator=g), 'noise': {'b8.conv0.noise_const': torch.randn(8, 8, generator=g),
                                                                  'b8.conv1.noise_const': torch.randn(8, 8, generator=g).numpy()}},
        'b': {'w': torch.randn(1, 14, 64, generator=g), 'b8.conv0.noise_const': torch.randn(8, 8, generator=g),
              'b8.conv1.noise_const': torch.randn(8, 8, generator=g)},                 # legacy layout: noise next to 'w'
        'c': torch.randn(1, 1, 64, generator=g),                                        # bare tensor
        'd': {'w': torch.randn(1, 14, 64, generator=g)},                                # dictionary without noise
    }
    path = os.path.join(tmp_path, 'brushes.pkl')
    with open(path, 'wb') as f:
        pickle.dump(styles, f)
    return path, styles


def test_seed_library_and_file_format(tmp_path):
    path = os.path.join(tmp_path, 'seeds.txt')
    with open(path, 'w') as f:
        f.write('# saved brushes\n594 0.1 0.2 0.3\n\nnot-a-seed 1 2\n17 1.0 2.0 3.0\n')
    zs, zdim = L.read_zs(path)
    assert zs == [594, 17] and zdim == 3
    lib = L.BrushLibrary.from_arg(path, z_dim=64)
    assert isinstance(lib, L.SeedBrushLibrary) and lib.get_style_ids() == ['17', '594'] and lib.zdim == 64
    o = GanBrushOptions()
    lib.set_style('594', o)
    assert o.style_id == '594' and o.style_ws is None
    assert np.array_equal(o.style_z.numpy(), np.random.RandomState(594).randn(1, 64))
    lib.set_interpolated_style('594', '17', 0.25, o)
    exp = np.random.RandomState(594).randn(1, 64) * 0.25 + np.random.RandomState(17).randn(1, 64) * 0.75
    assert np.array_equal(o.style_z.numpy(), exp) and o.style_id == '594_0.25__17'
    assert isinstance(L.BrushLibrary.from_arg('rand7'), L.RandomBrushLibrary)
    assert L.BrushLibrary.from_arg('3,5,9').get_style_ids() == ['3', '5', '9']
    assert len(L.BrushLibrary.from_arg('12').get_style_ids()) == 12
    with pytest.raises(ValueError):
        L.BrushLibrary.from_arg('3,x')


def test_w_library_styles_and_interpolation(tmp_path):
    path, styles = _wlib(tmp_path)
    lib = L.BrushLibrary.from_arg(path)
    assert isinstance(lib, L.WBrushLibrary) and lib.get_style_ids() == ['a', 'b', 'c', 'd']
    o = GanBrushOptions()
    lib.set_style('a', o)
    assert o.style_z is None and torch.equal(o.style_ws, styles['a']['w'])
    nb = o.custom_args['noise_buffers']
    assert set(nb) == {'b8.conv0.noise_const', 'b8.conv1.noise_const'} and all(torch.is_tensor(v) for v in nb.values())
    lib.set_style('b', o)
    assert set(o.custom_args['noise_buffers']) == {'b8.conv0.noise_const', 'b8.conv1.noise_const'}
    lib.set_style('c', o)
    assert o.custom_args['noise_buffers'] is None and o.style_ws.shape == (1, 1, 64)
    lib.set_style('d', o)
    assert o.custom_args['noise_buffers'] is None
    lib.set_interpolated_style('a', 'b', 0.3, o)
    assert torch.equal(o.style_ws, styles['a']['w'] * 0.3 + styles['b']['w'] * (1 - 0.3))
    k = 'b8.conv0.noise_const'
    assert torch.equal(o.custom_args['noise_buffers'][k], styles['a']['noise'][k] * 0.3 + styles['b'][k] * (1 - 0.3))
    assert o.style_id == 'a_0.30__b'


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference only exists in the build container')
def test_matches_reference_library(tmp_path):
    # stand-ins for two imports of forger.ui.brush that are absent here and unused numerically
    for name, attrs in {'skimage': {}, 'skimage.io': {'imread': None, 'imsave': None}, 'matplotlib': {}, 'matplotlib.pyplot': {}}.items():
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            sys.modules[name] = m
    sys.modules['skimage'].io = sys.modules['skimage.io']
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import thirdparty.stylegan2_ada_pytorch  # noqa: F401  (puts the stylegan root on sys.path)
    import forger.ui.library as RL
    import forger.ui.brush as RB
    path, _ = _wlib(tmp_path)
    ours, ref = L.WBrushLibrary.from_file(path), RL.WBrushLibrary.from_file(path)
    assert ours.get_style_ids() == ref.get_style_ids()
    # styles WITHOUT noise maps: the reference dereferences the None it stored itself (library.py:196-198) and raises; here
    # the w+ codes are interpolated and no noise buffers are passed on
    with pytest.raises(AttributeError):
        ref.set_interpolated_style('c', 'd', 0.5, RB.GanBrushOptions())
    o = GanBrushOptions()
    ours.set_interpolated_style('c', 'd', 0.5, o)
    assert o.style_ws.shape == (1, 14, 64) and not o.custom_args
    for args in (('a', 'b', 0.3), ('b', 'a', 0.9)):
        o, r = GanBrushOptions(), RB.GanBrushOptions()
        ours.set_interpolated_style(*args, o)
        ref.set_interpolated_style(*args, r)
        assert o.style_id == r.style_id and torch.equal(o.style_ws, r.style_ws)
        on, rn = o.custom_args.get('noise_buffers'), r.custom_args.get('noise_buffers')
        assert (on is None) == (rn is None)
        if on is not None:
            assert set(on) == set(rn) and all(torch.equal(on[k], rn[k]) for k in on)
    so, sr = L.SeedBrushLibrary([594, 17], 64), RL.SeedBrushLibrary([594, 17], 64)
    o, r = GanBrushOptions(), RB.GanBrushOptions()
    so.set_interpolated_style('594', '17', 0.4, o)
    sr.set_interpolated_style('594', '17', 0.4, r)
    assert o.style_id == r.style_id and torch.equal(o.style_z, r.style_z)
    ro, rr = L.RandomBrushLibrary(3, 64), RL.RandomBrushLibrary(3, 64)
    for _ in range(3):
        ro.set_style('rand0', o)
        rr.set_style('rand0', r)
        assert torch.equal(o.style_z, r.style_z)
    assert L.interp_style_id('a', 7, 0.5) == RL._interp_style_id('a', 7, 0.5)


def test_prepare_colors_owned_and_default_colour_range():
    """``GanBrushOptions.prepare_colors`` (brush.py:514-527): user colours override columns of the default table; the caller's table
    is cloned unless the caller says it owns it.  ``_unit_range_colors`` is ``(colors + 1) / 2`` (brush.py:770,913); on a CPU tensor
    it is the torch expression itself (the one-launch form needs a CUDA tensor)."""
    import torch
    from brushstroke_engine_b200 import engine as E
    colors = torch.tensor([[[-1.0, 0.0, 1.0], [0.5, -0.5, 0.25], [1.0, 1.0, -1.0]]]).repeat(4, 1, 1)
    unit = E._unit_range_colors(colors)
    assert torch.equal(unit, (colors + 1) / 2.0) and float(unit.min()) >= 0.0 and float(unit.max()) <= 1.0
    opts = E.GanBrushOptions()
    kept = unit.clone()
    out = opts.prepare_colors(unit)
    assert out is not unit and torch.equal(out, kept)
    assert opts.prepare_colors(unit, owned=True) is unit
    opts.color1 = torch.tensor([0.1, 0.2, 0.3])
    out = opts.prepare_colors(unit)
    assert torch.equal(unit, kept)                                  # the caller's table is untouched
    assert torch.equal(out[:, :, 1], torch.tensor([0.1, 0.2, 0.3]).expand(4, 3)) and torch.equal(out[:, :, 0], kept[:, :, 0])
    out2 = opts.prepare_colors(unit, owned=True)
    assert out2 is unit and torch.equal(out2, out)
