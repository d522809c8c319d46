"""GPU: nbe_blend_window_nhwc_bf16 against the torch statement of PaintingHelper's feature canvas (forger/ui/brush.py:190-242:
look-up of saved features, dirty-area alpha, core write-back) and BlendedFeatures.blend (forger/train/stitching.py:18-25)."""
import numpy as np
import pytest
import torch

from brushstroke_engine_b200 import _lib
from brushstroke_engine_b200.stylizer import dirty_area_alpha

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.mark.parametrize('R,C,pitch,cs,cm,with_scale', [(64, 128, 65, 128, 5, True), (32, 128, 33, 384, 2, True), (128, 128, 128, 128, 10, False),
                                                      (16, 64, 20, 64, 0, True)])
def test_blend_window_matches_torch(R, C, pitch, cs, cm, with_scale):
    g = torch.Generator().manual_seed(R + C + cm)
    n, FH, FW = 3, 3 * R, 4 * R
    base_alpha = dirty_area_alpha(R, max(1, 16 * R // 128), cm, DEV).float().contiguous()
    x = torch.zeros((n, R, pitch, cs), dtype=torch.bfloat16, device=DEV)
    x[:, :, :R, :] = torch.randn((n, R, R, cs), generator=g).to(DEV).to(torch.bfloat16)
    fcanvas = torch.randn((FH, FW, C), generator=g).to(DEV).to(torch.bfloat16)
    fmask = (torch.rand((FH, FW), generator=g) > 0.5).to(DEV).to(torch.uint8)
    fyx = torch.tensor([[0, 0], [R // 2, R + 3], [FH - R, FW - R]], dtype=torch.int32, device=DEV)      # disjoint windows
    scale = (torch.rand((n, C), generator=g) + 0.5).to(DEV) if with_scale else None
    x0, f0, m0 = x.clone(), fcanvas.clone(), fmask.clone()
    _lib.call('nbe_blend_window_nhwc_bf16', _lib.ptr(x), pitch, cs, R, C, _lib.ptr(fcanvas), _lib.ptr(fmask), FH, FW, _lib.ptr(fyx),
              _lib.ptr(base_alpha), cm, _lib.ptr(scale), n, _lib.stream())
    torch.cuda.synchronize()
    inner = torch.zeros((R, R), dtype=torch.bool, device=DEV)
    inner[cm:R - cm, cm:R - cm] = True
    for i in range(n):
        fy, fx = int(fyx[i, 0]), int(fyx[i, 1])
        m = m0[fy:fy + R, fx:fx + R] != 0
        alpha = torch.where(m, base_alpha, torch.ones((), device=DEV))
        a = (1 - alpha)[..., None]
        saved = f0[fy:fy + R, fx:fx + R].float()
        xb = (a * saved + (1 - a) * x0[i, :, :R, :C].float()).to(torch.bfloat16)
        update = ((base_alpha > 0.99) | (m & (base_alpha > 0))) & inner
        want_f = torch.where(update[..., None], xb, f0[fy:fy + R, fx:fx + R])
        assert torch.equal(fcanvas[fy:fy + R, fx:fx + R].view(torch.int16), want_f.view(torch.int16)), i
        assert torch.equal(fmask[fy:fy + R, fx:fx + R] != 0, m | update), i
        want_x = (xb.float() * scale[i]).to(torch.bfloat16) if with_scale else xb
        assert torch.equal(x[i, :, :R, :C].view(torch.int16), want_x.view(torch.int16)), i
        # gap columns and the channels beyond C are untouched
        assert torch.equal(x[i, :, R:, :].view(torch.int16), x0[i, :, R:, :].view(torch.int16))
        assert torch.equal(x[i, :, :R, C:].view(torch.int16), x0[i, :, :R, C:].view(torch.int16))
    # nothing outside the windows changed
    untouched = torch.ones((FH, FW), dtype=torch.bool, device=DEV)
    for i in range(n):
        fy, fx = int(fyx[i, 0]), int(fyx[i, 1])
        untouched[fy:fy + R, fx:fx + R] = False
    assert torch.equal(fcanvas[untouched].view(torch.int16), f0[untouched].view(torch.int16))
    assert torch.equal(fmask[untouched], m0[untouched])


def test_blend_window_validates_arguments():
    P = 4096
    with pytest.raises(RuntimeError, match='must divide 32'):
        _lib.call('nbe_blend_window_nhwc_bf16', P, 65, 128, 64, 96, P, P, 128, 128, P, P, 0, None, 1, None)
    with pytest.raises(RuntimeError, match='bad pitches'):
        _lib.call('nbe_blend_window_nhwc_bf16', P, 60, 128, 64, 128, P, P, 128, 128, P, P, 0, None, 1, None)
