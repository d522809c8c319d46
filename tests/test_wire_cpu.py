"""CPU: the wire codec of the interactive path (brushstroke_engine_b200.server) against a session recorded from the
reference's own DrawingWebSocketHandler (tests/golden/wire.npz, written by oracle/make_golden.py --only wire):
forger/ui/util.py:21-104 and the client's encoder / decoder forger/ui/js/main_controller.js:532-677."""
import json

import numpy as np
import pytest

from conftest import load_golden
from brushstroke_engine_b200 import server


def _messages(g):
    for i in range(int(g['n_msgs'])):
        raw = g[f'm{i}_in'].tobytes()
        outs = [(g[f'm{i}_out{j}'].tobytes(), bool(g[f'm{i}_out{j}_binary'])) for j in range(int(g[f'm{i}_nout']))]
        yield i, (raw if g[f'm{i}_in_binary'] else raw.decode()), outs


def test_requests_decode_and_reencode_bit_exactly():
    g = load_golden('wire')
    n = 0
    for i, msg, outs in _messages(g):
        if not isinstance(msg, bytes) or len(msg) < 100:
            continue
        meta, off = server.decode_render_request_metadata(msg)
        pm, stroke, canvas = server.binary_to_image_patches(msg, off)
        assert canvas is None and stroke.shape == (pm['height'], pm['width'], 4) and stroke.dtype == np.uint8
        again = server.encode_render_request(stroke, int(pm['x']), int(pm['y']), int(pm['crop_margin']),
                                             colors=[tuple(int(v) for v in c) for c in meta['colors']],
                                             debug=bool(meta['debug']), extra_data=int(meta['extra_data']))
        assert again == msg
        n += 1
    assert n == 8


def test_responses_decode_and_reencode_bit_exactly():
    g = load_golden('wire')
    n = 0
    for i, msg, outs in _messages(g):
        for payload, binary in outs:
            if not binary:
                json.loads(payload.decode())
                continue
            kind, meta, img = server.decode_render_response(payload)
            assert img.shape == (meta['height'], meta['width'], 4)
            assert server.int32_to_binary(kind) + server.image_patch_to_binary(img, meta['x'], meta['y']) == payload
            n += 1
    assert n == 8


def test_recorded_session_shape():
    """What the reference handler answers to what (pins the session logic the GPU test replays)."""
    g = load_golden('wire')
    assert int(g['open_nout']) == 2
    kinds = []
    for i, msg, outs in _messages(g):
        if isinstance(msg, bytes):
            kinds.append(('bin', len(outs)))
        else:
            kinds.append((json.loads(msg)['type'], len(outs)))
    assert kinds.count(('bin', 1)) == 8 and ('bin', 0) in kinds            # the truncated request is dropped silently
    assert ('set_brush', 1) in kinds and ('new_canvas', 0) in kinds and ('bogus', 0) in kinds
    # answer to the request with extra_data = 5 echoes it as the response type; crop margin moves the patch origin
    req = [m for _, m, _ in _messages(g) if isinstance(m, bytes) and len(m) > 100]
    outs = [o for _, m, o in _messages(g) if isinstance(m, bytes) and len(m) > 100]
    kind, meta, img = server.decode_render_response(outs[1][0][0])
    assert kind == 5 and (meta['x'], meta['y']) == (88 + 10, 0 + 10) and img.shape == (108, 108, 4)
    kind, meta, img = server.decode_render_response(outs[3][0][0])
    assert kind == 0 and (meta['x'], meta['y']) == (176, 88) and img.shape == (128, 128, 4)


def test_codec_rejects_bad_input():
    with pytest.raises(RuntimeError):
        server.image_patch_to_binary(np.zeros((8, 8, 4), dtype=np.float32), 0, 0)
    with pytest.raises(RuntimeError):
        server.encode_render_request(np.zeros((8, 8, 3), dtype=np.uint8), 0, 0)
    with pytest.raises(ValueError):
        server.binary_to_image_patches(b'\x00' * 10)
