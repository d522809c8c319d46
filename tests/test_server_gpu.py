"""GPU: the interactive path (server.DrawingSession / PaintingHelper / StrokeBatcher) against the session recorded from
the reference's DrawingWebSocketHandler + PaintingHelper (tests/golden/wire.npz; forger/ui/util.py:107-245,
forger/ui/brush.py:95-398): response headers exact, pixels within the image tolerance of the mode."""
import json
import re

import numpy as np
import pytest
import torch

from conftest import load_golden
from brushstroke_engine_b200 import params as P, server, synthetic
from brushstroke_engine_b200.engine import TriadPaintEngine

pytestmark = pytest.mark.gpu
DEV = 'cuda'


@pytest.fixture(scope='module')
def engines(bundles):
    cfg, ecfg, gp, ep = bundles
    return {m: TriadPaintEngine(gp, ep, DEV, mode=m, gen_cfg=cfg, enc_cfg=ecfg) for m in ('fp32', 'bf16')}


def _rgb(s):
    return np.array([int(v) for v in re.findall(r'\d+', s)])


def _check(ours, theirs, lsb):
    assert len(ours) == len(theirs)
    for (payload, binary), (ref, ref_binary) in zip(ours, theirs):
        assert binary == ref_binary
        if not binary:
            a, b = payload, json.loads(ref.decode())
            assert a['type'] == b['type']
            if a['type'] == 'brushinfo':
                assert a['data']['style_id'] == b['data']['style_id'] and a['data']['library_id'] == b['data']['library_id']
                assert np.abs(_rgb(a['data']['colors']) - _rgb(b['data']['colors'])).max() <= lsb
            else:
                assert a == b
            continue
        assert payload[:20] == ref[:20]                                   # type, width, height, x, y
        _, _, img = server.decode_render_response(payload)
        _, _, want = server.decode_render_response(ref)
        d = np.abs(img.astype(np.int32) - want.astype(np.int32))
        assert d.max() <= lsb, d.max()


@pytest.mark.parametrize('mode,lsb', [('fp32', 1), ('bf16', 3)])
def test_session_replays_the_reference_transcript(engines, mode, lsb):
    g = load_golden('wire')
    eng = engines[mode]
    eng.set_render_mode('clear')
    sess = server.DrawingSession(eng, style_seed=3)
    opened = [(g[f'open_out{j}'].tobytes(), bool(g[f'open_out{j}_binary'])) for j in range(int(g['open_nout']))]
    _check(sess.open(), opened, lsb)
    for i in range(int(g['n_msgs'])):
        raw = g[f'm{i}_in'].tobytes()
        msg = raw if g[f'm{i}_in_binary'] else raw.decode()
        outs = [(g[f'm{i}_out{j}'].tobytes(), bool(g[f'm{i}_out{j}_binary'])) for j in range(int(g[f'm{i}_nout']))]
        _check(sess.on_message(msg), outs, lsb)
    sess.close()
    eng.set_render_mode('clear')


def _stroke(seed):
    geo = synthetic.synthetic_patch(128, seed=seed, radius=5)[0, 0]
    rgba = np.zeros((128, 128, 4), dtype=np.uint8)
    rgba[..., 3] = np.round((1 - geo) * 255).astype(np.uint8)
    return rgba


def test_batched_sessions_equal_one_at_a_time(engines):
    """Six concurrent sessions (different brushes, colours, positions on / off, two of them with feature blending, one
    of those sending two overlapping strokes) rendered by ONE flush equal the same requests rendered session by session."""
    eng = engines['bf16']
    eng.set_render_mode('clear')

    def make(batcher):
        ss = [server.DrawingSession(eng, style_seed=100 + k, batcher=batcher) for k in range(6)]
        for k, s in enumerate(ss):
            s.on_message(json.dumps({'type': 'set_brush', 'seed': 500 + 7 * k}))
            s.on_message(json.dumps({'type': 'set_option', 'option': 'positions', 'value': k != 3}))
            s.on_message(json.dumps({'type': 'new_canvas', 'rows': 256, 'cols': 384, 'feature_blending': 2 if k in (1, 4) else 0}))
        return ss

    def requests():
        out = []
        for k in range(6):
            cols = [(1, 20 * k, 255 - 30 * k, 7)] if k % 2 else []
            out.append((k, server.encode_render_request(_stroke(40 + k), 16 * k + (k % 2), 10 * k, 10, colors=cols, extra_data=k)))
        out.append((1, server.encode_render_request(_stroke(50), 60, 40, 10)))           # overlaps session 1's first stroke
        out.append((4, server.encode_render_request(_stroke(51), 100, 70, 10)))
        return out

    solo = make(None)
    want = {k: [] for k in range(6)}
    for k, m in requests():
        want[k] += solo[k].on_message(m)
    batcher = server.StrokeBatcher(eng)
    ss = make(batcher)
    for k, m in requests():
        assert ss[k].on_message(m) == []
    got = server.DrawingSession.flush_all(batcher, ss)
    # {positions, no blending: 3 sessions} + {no positions: 1} + {blending: wave 0 = first strokes of sessions 1 and 4, wave 1 = second}
    assert batcher.forwards == 4
    for k in range(6):
        assert [p for p, _ in got[ss[k]]] == [p for p, _ in want[k]], k
    for s in solo + ss:
        s.close()


def test_feature_pool_bands():
    cfg, ecfg = P.GeneratorConfig(), P.EncoderConfig()
    eng = TriadPaintEngine(P.init_generator_params(cfg, 0, 0.1), P.init_encoder_params(ecfg, 1, 0.1), DEV, mode='bf16')
    eng.feature_pool_bytes = 256 << 20
    a, b = server.PaintingHelper(eng, 1), server.PaintingHelper(eng, 2)
    a.make_new_canvas(256, 256, feature_blending=2)
    b.make_new_canvas(512, 300, feature_blending=2)
    pool = a._pool
    assert pool is b._pool and a._band[0] == 0 and b._band[0] == 128 + 64
    pool.fmask[a._band[0]:a._band[0] + 10].fill_(1)
    a.make_new_canvas(128, 128, feature_blending=2)                         # released and re-acquired: zeroed again
    assert a._band[0] == 0 and int(pool.fmask[:10].sum()) == 0
    a.close(); b.close()
    assert pool._free == [(0, pool.fcanvas.shape[0])]
    with pytest.raises(RuntimeError, match='starts outside'):
        a.make_new_canvas(128, 128, feature_blending=2)
        a.render_stroke(_stroke(1), None, a.default_brush_options(), {'x': 500, 'y': 0, 'crop_margin': 0})
