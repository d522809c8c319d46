"""CPU: the C-ABI library builds, loads and exports exactly what include/nbe_b200.h declares."""
import ctypes
import os
import re

import pytest

from conftest import REPO
from brushstroke_engine_b200 import _lib, build


def header_functions():
    src = open(os.path.join(REPO, 'include', 'nbe_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    out = {}
    for m in re.finditer(r'\b(?:int|int64_t|const char\*)\s+(nbe_\w+)\s*\(([^;{]*?)\)\s*;', src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ('', 'void') else len([a for a in args.split(',') if a.strip()])
        out[m.group(1)] = n
    return out


@pytest.fixture(scope='module')
def lib():
    build.build()
    return ctypes.CDLL(_lib.LIB_PATH)


def test_every_declared_symbol_is_exported(lib):
    decl = header_functions()
    assert len(decl) >= 20
    for name in decl:
        assert hasattr(lib, name), f'{name} declared in include/nbe_b200.h but not exported'


def test_binding_matches_header(lib):
    decl = header_functions()
    for name, argtypes in _lib._SIGNATURES.items():
        assert name in decl, f'{name} bound in _lib.py but not declared in the header'
        assert decl[name] == len(argtypes), f'{name}: header has {decl[name]} parameters, binding has {len(argtypes)}'
    assert set(decl) == set(_lib.exported_symbols())


def test_abi_version_and_error_string(lib):
    lib.nbe_abi_version.restype = ctypes.c_int
    lib.nbe_last_error.restype = ctypes.c_char_p
    assert lib.nbe_abi_version() == 1
    assert isinstance(lib.nbe_last_error(), bytes)


def test_invalid_arguments_are_reported_not_thrown():
    """Validation happens before any CUDA call, so this is safe without a GPU."""
    L = _lib.load()
    st = L.nbe_bias_act(None, None, None, 16, 0, 1, 3, 0.2, 1.0, -1.0, 0, None)
    assert st == -1 and b'null' in L.nbe_last_error()
    st = L.nbe_bias_act(ctypes.c_void_p(16), None, ctypes.c_void_p(16), 16, 0, 1, 42, 0.2, 1.0, -1.0, 0, None)
    assert st == -1 and b'activation' in L.nbe_last_error()
    with pytest.raises(RuntimeError, match='nbe_upfirdn2d failed'):
        _lib.call('nbe_upfirdn2d', 16, 16, 16, 1, 1, 4, 4, 16, 16, 4, 1, 4, 4, 16, 16, 4, 1,
                  4, 4, 0, 1, 1, 1, 0, 0, 0, 0, 0, 1.0, 0, None)


def test_ops_refuse_cpu_tensors():
    import torch
    from brushstroke_engine_b200 import bias_act, upfirdn2d, modconv
    x = torch.zeros(1, 2, 4, 4)
    with pytest.raises(RuntimeError, match='CUDA'):
        bias_act.bias_act(x, act='lrelu')
    with pytest.raises(RuntimeError, match='CUDA'):
        upfirdn2d.upfirdn2d(x, upfirdn2d.setup_filter([1, 3, 3, 1]))
    with pytest.raises(RuntimeError, match='CUDA'):
        modconv.modulated_conv2d(x, torch.zeros(3, 2, 3, 3), torch.ones(1, 2))
    with pytest.raises(RuntimeError, match="impl='cuda'"):
        bias_act.bias_act(x, impl='ref')


def test_tensor_core_entry_points_validate_before_launching():
    """The tcgen05 / canvas entry points reject what they cannot run with a message naming the constraint; all of these
    checks sit in front of the first CUDA call (dummy non-null pointers, no GPU needed)."""
    P = 4096                                                             # any non-null, 16-byte aligned "pointer"
    cases = [
        ('Cout must be a multiple of 128', 'nbe_conv3x3_flat_bf16',
         (P, P, P, 1, 8, 8, 64, 64, 9, 0, 96, 96, 8, 64, None, None, 0, 0.0, None, 0.2, 1.0, -1.0, None, None)),
        ('input pitch', 'nbe_conv3x3_flat_bf16',
         (P, P, P, 1, 8, 8, 64, 64, 8, 0, 128, 128, 8, 64, None, None, 0, 0.0, None, 0.2, 1.0, -1.0, None, None)),
        ('16-byte aligned', 'nbe_conv3x3_flat_bf16',
         (P + 2, P, P, 1, 8, 8, 64, 64, 9, 0, 128, 128, 8, 64, None, None, 0, 0.0, None, 0.2, 1.0, -1.0, None, None)),
        ('H, W must be even', 'nbe_conv3x3s2_flat_bf16',
         (P, P, P, 1, 7, 8, 64, 128, 128, 4, 16, None, 0.01, 1.0, -1.0, None, None)),
        ('Cin must be a multiple of 64', 'nbe_conv3x3s2_flat_bf16',
         (P, P, P, 1, 8, 8, 48, 128, 128, 4, 16, None, 0.01, 1.0, -1.0, None, None)),
        ('zero gap column', 'nbe_convT3x3s2_flat_bf16',
         (P, P, P, 1, 8, 8, 128, 128, 8, 128, 128, 18, 18 * 18, None, None)),
        ('bad output pitches', 'nbe_convT3x3s2_flat_bf16',
         (P, P, P, 1, 8, 8, 128, 128, 9, 128, 128, 16, 16 * 16, None, None)),
        ('does not match input', 'nbe_fir_act_nhwc_bf16',
         (P, P, P, 1, 16, 16, 128, 16, 16, 1, 128, 18, 18 * 18, 128, 16, 256, 4.0, None, None, 0, 0.0, None, 0.2, 1.0, -1.0, None, None)),
        ('unknown render mode', 'nbe_canvas_composite', (P, P, P, 0, P, 7, P, None, 1, 8, 8, 0, None)),
        ('no output', 'nbe_canvas_composite', (P, P, P, 0, P, 0, None, None, 1, 8, 8, 0, None)),
        ('too many channels', 'nbe_torgb_canvas', (P, 0, 0, P, P, P, P, 256.0, P, P, P, P, 1, 4096, 8, 8, None)),
    ]
    for needle, name, args in cases:
        with pytest.raises(RuntimeError, match=name) as ei:
            _lib.call(name, *args)
        assert needle in str(ei.value), (name, needle, str(ei.value))
