"""Synthetic inputs for the benchmark configs of SURVEY.md section 8(d).

There is no network and no dataset on the GPU box, so guidance drawings are
generated: a white uint8 canvas with K random poly-lines stamped as discs
(value 0 = stroke), radii taken from the reference's bundled fixtures
(forger/images/spline_patches_curated/*_rad{016,025}.png use radii 16 and 25;
thin lines added so that fine geometry is exercised too).
"""
from __future__ import annotations

import numpy as np


def _stamp_disc(img: np.ndarray, cy: float, cx: float, r: int) -> None:
    H, W = img.shape
    y0, y1 = max(int(cy) - r - 1, 0), min(int(cy) + r + 2, H)
    x0, x1 = max(int(cx) - r - 1, 0), min(int(cx) + r + 2, W)
    if y0 >= y1 or x0 >= x1:
        return
    yy, xx = np.mgrid[y0:y1, x0:x1]
    img[y0:y1, x0:x1][(yy - cy) ** 2 + (xx - cx) ** 2 <= r * r] = 0


def synthetic_guidance(height: int, width: int, num_lines: int = 64, seed: int = 0,
                       radii=(1, 3, 9, 16, 25)) -> np.ndarray:
    """-> [H, W, 1] uint8, 255 = background, 0 = stroke (black on white, like the
    files ``paint_image_main._read_any_geo`` produces after Otsu thresholding)."""
    rng = np.random.default_rng(seed)
    img = np.full((height, width), 255, dtype=np.uint8)
    for _ in range(num_lines):
        r = int(radii[rng.integers(0, len(radii))])
        npts = int(rng.integers(2, 6))
        pts = np.stack([rng.uniform(0, height, npts), rng.uniform(0, width, npts)], axis=1)
        for a, b in zip(pts[:-1], pts[1:]):
            n = int(max(abs(b[0] - a[0]), abs(b[1] - a[1])) / max(r * 0.5, 1.0)) + 2
            for t in np.linspace(0.0, 1.0, n):
                _stamp_disc(img, a[0] + t * (b[0] - a[0]), a[1] + t * (b[1] - a[1]), r)
    return img[:, :, None]


def synthetic_patch(width: int = 128, seed: int = 0, radius: int = 8) -> np.ndarray:
    """One 128x128 geometry patch (a cross-like pair of strokes), float32 in
    [0, 1], 0 = stroke, shape [1, 1, W, W] -- the engine's ``geom`` convention
    (forger/ui/brush.py:672-681)."""
    rng = np.random.default_rng(seed)
    img = np.full((width, width), 255, dtype=np.uint8)
    for _ in range(2):
        a = rng.uniform(0, width, 2)
        b = rng.uniform(0, width, 2)
        for t in np.linspace(0.0, 1.0, 4 * width // max(radius, 1)):
            _stamp_disc(img, a[0] + t * (b[0] - a[0]), a[1] + t * (b[1] - a[1]), radius)
    return (img.astype(np.float32) / 255.0)[None, None]
