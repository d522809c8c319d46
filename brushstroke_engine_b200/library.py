"""Brush libraries: where the styles that drive the hot path come from.

Host-side plumbing with the reference's class / method names (``forger/ui/library.py:72-257``) so that code written
against ``BrushLibrary.from_arg(...)`` / ``set_style`` / ``set_interpolated_style`` keeps working with
``brushstroke_engine_b200.engine.GanBrushOptions``:

* ``SeedBrushLibrary``  -- integer seeds; z = ``RandomState(seed).randn(1, z_dim)`` (library.py:222-225); a saved-seed text
  file has one style per line, ``<int seed> <z0> <z1> ...``, ``#`` comments, only the seed is used (library.py:49-65);
* ``WBrushLibrary``     -- pickled ``style_id -> tensor`` or ``style_id -> {'w': tensor[1, num_ws or 1, w_dim], 'noise': {layer
  name: tensor[res, res]}}`` projections (library.py:146-186); the noise maps go to the generator as ``noise_buffers``;
* ``RandomBrushLibrary`` -- ``rand<N>``: a fresh random z per request.

Interpolation is linear in z (seeds) or in w+ and every noise buffer (projections), weight ``alpha`` on the FIRST style
(library.py:188-202, 227-234).  Icon caches (``*.icons.zip``) and dynamic icon rendering are UI features and not covered.
"""
from __future__ import annotations

import os
import pickle
import random
import re
from typing import Dict, List, Optional

import numpy as np
import torch

from .engine import GanBrushOptions


def read_zs(saved_file: str):
    """-> (list of int seeds, z_dim) from a saved-seed text file (library.py:49-65); ([], 0) when nothing parses."""
    zs: List[int] = []
    zdim = 0
    if not os.path.isfile(saved_file):
        return zs, zdim
    with open(saved_file) as f:
        for line in f:
            fields = line.strip().split()
            if not fields or fields[0].startswith('#'):
                continue
            try:
                seed = int(fields[0])
            except ValueError:
                continue                                            # the reference logs and skips such lines
            zdim = len(fields) - 1
            zs.append(seed)
    return zs, zdim


def interp_style_id(style_id1, style_id2, alpha: float) -> str:
    """library.py:68-69."""
    return '%s_%0.2f__%s' % (str(style_id1), alpha, str(style_id2))


class BrushLibrary:
    @staticmethod
    def from_arg(arg_val: str, z_dim: int = 64) -> 'BrushLibrary':
        """A file (w+ pickle or seed list), ``rand<N>``, ``N`` (N shuffled seeds) or ``s0,s1,...`` (library.py:73-98)."""
        if os.path.isfile(arg_val):
            return BrushLibrary.from_file(arg_val, z_dim=z_dim)
        m = re.match(r'^rand(\d+)$', arg_val)
        if m is not None:
            return RandomBrushLibrary(int(m.group(1)), zdim=z_dim)
        try:
            values = [int(x) for x in arg_val.split(',')]
        except ValueError as e:
            raise ValueError(f'style seeds must be comma-separated ints, got: {arg_val}') from e
        if len(values) == 1:
            seeds = list(range(0, max(10000, values[0])))
            random.shuffle(seeds)
            return SeedBrushLibrary(seeds[:values[0]], z_dim)
        return SeedBrushLibrary(values, z_dim)

    @staticmethod
    def from_file(fname: str, z_dim: int = 64) -> 'BrushLibrary':
        """A w+ pickle if it unpickles, else a seed list (library.py:100-110)."""
        try:
            return WBrushLibrary.from_file(fname)
        except Exception:
            return SeedBrushLibrary.from_file(fname, z_dim=z_dim)

    def get_style_ids(self) -> List[str]:
        raise NotImplementedError

    def set_style(self, style_id, brush_options: GanBrushOptions) -> None:
        raise NotImplementedError

    def set_interpolated_style(self, style_id1, style_id2, alpha: float, brush_options: GanBrushOptions) -> None:
        raise NotImplementedError


class WBrushLibrary(BrushLibrary):
    def __init__(self, styles_dict: Dict):
        self.styles = styles_dict

    @staticmethod
    def from_file(fname: str) -> 'WBrushLibrary':
        styles: Dict = {}
        if os.path.isfile(fname):
            with open(fname, 'rb') as f:
                styles = pickle.load(f)
            if not isinstance(styles, dict):
                raise ValueError(f'{fname}: not a style dictionary')
        return WBrushLibrary(styles)

    def get_style_ids(self):
        return sorted(self.styles.keys())

    def set_style(self, style_id, brush_options):
        info = self.styles[style_id]
        noise: Optional[dict] = None
        if isinstance(info, dict):
            w = info['w']
            noise = info['noise'] if 'noise' in info else {k: v for k, v in info.items() if k != 'w'}
            if not noise:
                noise = None                                       # dictionary style without noise maps (library.py:173-175)
        else:
            w = info
        if noise is not None:
            noise = {k: (v if torch.is_tensor(v) else torch.from_numpy(np.asarray(v))) for k, v in noise.items()}
        brush_options.set_style_w(w, style_id=style_id, custom_args={'noise_buffers': noise})

    def set_interpolated_style(self, style_id1, style_id2, alpha, brush_options):
        o1, o2 = GanBrushOptions(), GanBrushOptions()
        self.set_style(style_id1, o1)
        self.set_style(style_id2, o2)
        w = o1.style_ws * alpha + o2.style_ws * (1 - alpha)
        custom_args = None
        n1, n2 = o1.custom_args.get('noise_buffers'), o2.custom_args.get('noise_buffers')
        if n1 is not None and n2 is not None:
            custom_args = {'noise_buffers': {k: v * alpha + n2[k] * (1 - alpha) for k, v in n1.items()}}
        brush_options.set_style_w(w, style_id=interp_style_id(style_id1, style_id2, alpha), custom_args=custom_args)


class SeedBrushLibrary(BrushLibrary):
    def __init__(self, seeds_list, zdim: int):
        self.zs = list(seeds_list)
        self.zdim = zdim

    @staticmethod
    def from_file(fname: str, z_dim: Optional[int] = None) -> 'SeedBrushLibrary':
        zs, zdim = read_zs(fname)
        return SeedBrushLibrary(zs, z_dim if z_dim is not None else zdim)

    def get_style_ids(self):
        return sorted(str(x) for x in self.zs)

    def set_style(self, style_id, brush_options):
        seed = int(style_id)
        brush_options.set_style(torch.from_numpy(np.random.RandomState(seed=seed).randn(1, self.zdim)), style_id=style_id)

    def set_interpolated_style(self, style_id1, style_id2, alpha, brush_options):
        o1, o2 = GanBrushOptions(), GanBrushOptions()
        self.set_style(style_id1, o1)
        self.set_style(style_id2, o2)
        brush_options.set_style(o1.style_z * alpha + o2.style_z * (1 - alpha), style_id=interp_style_id(style_id1, style_id2, alpha))


class RandomBrushLibrary(BrushLibrary):
    """``rand<N>``: N anonymous styles, a fresh z per request from one seeded stream.  The reference draws it with
    ``forger.metrics.util.RandomState(0).random_tensor`` = ``torch.rand`` (uniform, float32) from a ``torch.Generator``
    seeded with ``seed + 1`` (library.py:237-257, metrics/util.py:77-89); reproduced as is."""

    def __init__(self, num: int, zdim: int, generator: Optional[torch.Generator] = None):
        self.num = num
        self.zdim = zdim
        if generator is None:
            generator = torch.Generator()
            generator.manual_seed(0 + 1)
        self.generator = generator

    def get_style_ids(self):
        return ['rand' + str(x) for x in range(self.num)]

    def set_style(self, style_id, brush_options):
        brush_options.set_style(torch.rand((1, self.zdim), dtype=torch.float32, generator=self.generator))

    def set_interpolated_style(self, style_id1, style_id2, alpha, brush_options):
        self.set_style(style_id1, brush_options)
