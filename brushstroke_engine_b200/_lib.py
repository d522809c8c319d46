"""ctypes binding of libnbe_b200.so (the C ABI declared in include/nbe_b200.h).

The product path has no CPU or PyTorch fallback: if the shared library is
missing or a call fails, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_void_p

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libnbe_b200.so')

F32, F16, BF16, F64 = 0, 1, 2, 3
DTYPE_CODE = {torch.float32: F32, torch.float16: F16, torch.bfloat16: BF16, torch.float64: F64}

_P, _I, _L, _F = c_void_p, c_int, c_int64, c_float

# name -> argtypes  (order = include/nbe_b200.h)
_SIGNATURES = {
    'nbe_bias_act': [_P, _P, _P, _L, _L, _L, _I, _F, _F, _F, _I, _P],
    'nbe_upfirdn2d': [_P, _P, _P, _I, _I, _I, _I, _L, _L, _L, _L, _I, _I, _L, _L, _L, _L,
                      _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _F, _I, _P],
    'nbe_conv2d_f32': [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _L, _F, _P, _I, _F, _F, _F, _P],
    'nbe_weight_sqsum_f32': [_P, _P, _I, _I, _I, _P],
    'nbe_demod_coefs_f32': [_P, _P, _P, _I, _I, _I, _P],
    'nbe_fc_f32': [_P, _I, _P, _P, _P, _I, _I, _I, _L, _L, _F, _F, _I, _F, _F, _I, _P],
    'nbe_shifted_noise_f32': [_P, _P, _P, _P, _I, _I, _I, _P],
    'nbe_styles_demod_f32': [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    'nbe_styles_demod_input_f32': [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _I, _I, _I, _P],
    'nbe_shifted_noise_all_f32': [_P, _I, _I, _I, _P, _P, _P, _P, _P],
    'nbe_pack_nhwc_bf16': [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P],
    'nbe_unpack_nchw_f32': [_P, _P, _I, _I, _I, _I, _I, _P],
    'nbe_upsample2x_nhwc_bf16': [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    'nbe_upsample2x_nhwc_bf16_ex': [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    'nbe_prepare_weights_bf16': [_P, _P, _I, _I, _I, _I, _P],
    'nbe_conv_tc_bf16': [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _L, _F, _P, _F, _F, _F, _P, _P],
    'nbe_conv_tc_bf16_ex': [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _L, _L, _P, _P, _L, _F, _P, _F, _F, _F, _P, _P],
    'nbe_conv3x3_flat_bf16': [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _L, _L, _P, _P, _L, _F, _P, _F, _F, _F, _P, _P],
    'nbe_reflect_pad_nchw_f32': [_P, _P, _L, _I, _I, _I, _I, _P],
    'nbe_affine_nhwc_bf16': [_P, _I, _L, _L, _P, _I, _L, _L, _I, _I, _I, _I, _P, _P, _P, _P],
    'nbe_affine_nchw_f32': [_P, _P, _I, _I, _I, _P, _P, _P],
    'nbe_count_stroke_pixels': [_P, _I, _I, _I, _I, _I, _I, _P, _P],
    'nbe_torgb_canvas': [_P, _I, _I, _P, _P, _P, _P, _F, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    'nbe_canvas_composite': [_P, _P, _P, _L, _P, _I, _P, _P, _I, _I, _I, _I, _P],
    'nbe_conv3x3s2_flat_bf16': [_P, _P, _P, _I, _I, _I, _I, _I, _I, _L, _L, _P, _F, _F, _F, _P, _P],
    'nbe_convT3x3s2_flat_bf16': [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _L, _L, _P, _P],
    'nbe_fir_act_nhwc_bf16': [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _L, _L, _I, _L, _L, _F, _P, _P, _L, _F, _P, _F, _F, _F, _P, _P],
    'nbe_conv_tc_bf16_torgb': [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _L, _F, _P, _F, _F, _F, _P, _P, _P, _P, _F, _P, _P, _I, _P],
    'nbe_enc_conv7x7_bf16': [_P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _I, _P],
    'nbe_enc_conv7x7_tc_bf16': [_P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _I, _P],
    'nbe_enc_conv7x7_toeplitz_weights': [_P, _P, _I, _P],
    'nbe_enc_conv7x7_toeplitz_bf16': [_P, _P, _P, _P, _P, _L, _I, _I, _I, _I, _I, _F, _I, _P],
    'nbe_reflect_border_nhwc_bf16': [_P, _I, _I, _I, _I, _I, _P],
    'nbe_bilinear2x_pad_nhwc_bf16': [_P, _P, _I, _I, _I, _I, _I, _I, _P],
    'nbe_torgb_triad': [_P, _I, _I, _P, _P, _P, _P, _F, _P, _P, _I, _I, _I, _I, _P],
    'nbe_triad_composite': [_P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _P],
    'nbe_gather_geom_patches': [_P, _I, _I, _P, _P, _I, _I, _P],
    'nbe_tile_owner_map': [_P, _I, _I, _P, _I, _I, _P],
    'nbe_place_tiles': [_P, _P, _P, _I, _I, _P, _P, _I, _I, _P],
    'nbe_blend_window_nhwc_bf16': [_P, _I, _I, _I, _I, _P, _P, _I, _I, _P, _P, _I, _P, _I, _P],
    'nbe_blend_features': [_P, _P, _P, _L, _I, _I, _I, _I, _I, _I, _P],
    'nbe_up_layer_fused_bf16': [_P, _P, _P, _P, _P, _L, _I, _I, _I, _I, _I, _I, _I, _I, _L, _L, _F, _P, _P, _L, _F, _P, _F, _F, _F, _P, _P],
    'nbe_modulated_conv2d': [_P, _I, _P, _P, _P, _L, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _I, _I, _P, _L, _P],
}

# entry points that do not return a status: name -> (restype, argtypes)
_OTHER = {
    'nbe_modulated_conv2d_workspace': (c_int64, [_I, _I, _I, _I, _I, _I, _I, _I, _I]),
    'nbe_up_layer_fused_scratch_bytes': (c_int64, [_I]),
    'nbe_enc_conv7x7_toeplitz_scratch_bytes': (c_int64, [_I, _I, _I]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f'{LIB_PATH} is missing: build it with `python -m brushstroke_engine_b200.build` '
            '(there is no CPU / PyTorch fallback for the NeuBE hot path).')
    lib = ctypes.CDLL(LIB_PATH)
    lib.nbe_abi_version.restype = c_int
    lib.nbe_last_error.restype = c_char_p
    lib.nbe_launch_count.restype = c_int64
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = c_int
    for name, (restype, argtypes) in _OTHER.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    if lib.nbe_abi_version() != 1:
        raise RuntimeError(f'libnbe_b200.so ABI version {lib.nbe_abi_version()} != 1; rebuild it')
    _lib = lib
    return lib


def exported_symbols():
    return ['nbe_abi_version', 'nbe_last_error', 'nbe_launch_count'] + list(_SIGNATURES.keys()) + list(_OTHER.keys())


def launch_count() -> int:
    return int(load().nbe_launch_count())


def call(name: str, *args) -> None:
    """Invoke one entry point; non-zero status -> RuntimeError with the library's message."""
    lib = load()
    status = getattr(lib, name)(*args)
    if status != 0:
        msg = lib.nbe_last_error().decode('utf-8', 'replace')
        raise RuntimeError(f'{name} failed ({status}): {msg}')


# ---- profiler ranges --------------------------------------------------------------------------------------------------
# The reference marks its hot functions with `misc.profiled_function` / `record_function` (SG2/torch_utils/misc.py:98-103;
# names: 'modulated_conv2d', 'normalize_2nd_moment', '_bias_act_ref' -> 'bias_act', '_upfirdn2d_ref' -> 'upfirdn2d', 'input',
# 'broadcast', 'truncate', 'split_ws').  With NBE_NVTX=1 the same names appear as NVTX ranges around the corresponding launches
# (visible in nsys / ncu --nvtx); off by default: a range costs ~1 us of host time per call.
NVTX = os.environ.get('NBE_NVTX', '0') not in ('', '0')


class nvtx_range:
    __slots__ = ('name',)

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if NVTX:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if NVTX:
            torch.cuda.nvtx.range_pop()
        return False


def profiled(name: str):
    """Decorator form of ``nvtx_range`` (the reference's ``misc.profiled_function``)."""
    def deco(fn):
        if not NVTX:
            return fn
        import functools

        @functools.wraps(fn)
        def wrapped(*a, **k):
            torch.cuda.nvtx.range_push(name)
            try:
                return fn(*a, **k)
            finally:
                torch.cuda.nvtx.range_pop()
        return wrapped
    return deco


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not isinstance(t, torch.Tensor) or t.device.type != 'cuda':
        raise RuntimeError(f'{what}: expected a CUDA tensor, got {type(t).__name__} on '
                           f'{getattr(t, "device", None)} (no CPU fallback on this path)')
