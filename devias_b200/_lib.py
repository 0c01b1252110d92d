"""ctypes binding of the C-ABI shared library (include/devias_b200.h).

The library is built in-tree (devias_b200/csrc/libdevias_b200.so) by `__graft_entry__.build()` or
`make -C devias_b200/csrc`.  There is deliberately NO fallback: if the library is missing or a call
fails, a RuntimeError is raised (BASELINE.json north_star: "no CPU fallback").
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import c_char_p, c_float, c_int, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, 'csrc')
# DEVIAS_B200_LIB: developer switch for A/B builds of the same sources (e.g. -DDV_DEBUG_SPIN); default = the in-tree library
LIB_PATH = os.environ.get('DEVIAS_B200_LIB') or os.path.join(CSRC, 'libdevias_b200.so')

_lib = None


def build(verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(['make', '-C', CSRC, '-j8'], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('devias_b200: building the CUDA library failed:\n' + r.stdout[-4000:] + r.stderr[-4000:])
    if verbose:
        print(r.stdout[-2000:])
    return LIB_PATH


_P = c_void_p
_SIGS = {
    'devias_abi_version': (c_int, []),
    'devias_last_error': (c_char_p, []),
    'devias_launch_count': (c_int64, []),
    'devias_profile_begin': (c_int, []),
    'devias_profile_begin_capture': (c_int, []),
    'devias_profile_pause': (c_int, []),
    'devias_profile_end': (c_int, [c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                   ctypes.POINTER(c_int64)]),
    'devias_gemm_bf16': (c_int, [_P, c_int64, c_int, _P, c_int64, c_int, c_int, c_int, c_int, c_int, _P, c_int64, _P,
                                 c_int64, _P, _P, c_int64, c_int, _P, c_int, c_int, _P]),
    'devias_flash_attn_fwd': (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_float, _P]),
    'devias_flash_attn_bwd': (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_float, _P]),
    'devias_slot_stream_fwd': (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_float, _P]),
    'devias_slot_stream_fwd_bf16': (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_float, _P]),
    'devias_slot_stream_bwd': (c_int, [_P] * 11 + [c_int] + [_P] * 3 + [c_int, c_int, c_int, c_int, _P]),
    'devias_slot_stream_bwd_bf16': (c_int, [_P] * 11 + [c_int] + [_P] * 3 + [c_int, c_int, c_int, c_int, _P]),
    'devias_layernorm_fwd': (c_int, [_P, _P, _P, _P, c_int, _P, _P, c_int, c_int, c_float, _P]),
    'devias_layernorm_bwd': (c_int, [_P, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P]),
    'devias_colsum_bf16': (c_int, [_P, c_int64, c_int, c_int, _P, _P]),
    'devias_cast_f32_bf16': (c_int, [_P, _P, c_int64, _P]),
    'devias_cast_bf16_f32': (c_int, [_P, _P, c_int64, _P]),
    'devias_scale_rows_cast': (c_int, [_P, _P, c_int, c_int, _P, c_int, _P]),
    'devias_patch_embed_fwd': (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    'devias_patchify': (c_int, [_P, c_int, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    'devias_slot_fold_fwd': (c_int, [_P, _P, _P, c_float, _P, _P, _P, c_int, c_int, _P]),
    'devias_slot_fold_bwd': (c_int, [_P, _P, _P, c_float, _P, _P, _P, _P, _P, _P, c_int, c_int, _P]),
    'devias_slot_ctx_fwd': (c_int, [_P, _P, _P, _P, _P, c_float, _P, c_int, c_int, _P]),
    'devias_slot_ctx_bwd': (c_int, [_P, _P, _P, _P, _P, _P, c_float, _P, _P, _P, _P, _P, c_int, c_int, _P]),
    'devias_slot_select': (c_int, [_P, c_int64, c_int, c_int, c_int, c_int, _P, _P, _P]),
    'devias_debug_token_stream': (c_int, [_P, c_int, c_int, c_int, _P, _P]),
    'devias_skinny_nt': (c_int, [_P, _P, _P, c_int64, _P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    'devias_skinny_nn': (c_int, [_P, _P, _P, c_int64, _P, _P, c_int, c_int, c_int, c_int, _P]),
    'devias_train_loss_fwd': (c_int, [_P] * 9 + [c_int] * 9 + [c_float] * 3 + [_P, _P, _P]),
    'devias_train_loss_bwd': (c_int, [_P] * 9 + [c_int] * 9 + [c_float] * 3 + [_P] * 5 + [_P]),
    'devias_sumsq_f32': (c_int, [_P, c_int64, _P, _P]),
    'devias_adamw_arena': (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int, _P, _P, c_int64, c_int, _P, _P]),
    'devias_skinny_outer': (c_int, [_P, _P, _P, _P, _P, c_int64, _P, c_int64, c_int, c_int, c_int, c_int, c_int, _P]),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(f'devias_b200: {LIB_PATH} is missing -- run __graft_entry__.build() '
                               f'(or make -C devias_b200/csrc). There is no CPU/PyTorch fallback.')
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)  # AttributeError if the .so is stale: fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def exported_symbols():
    return list(_SIGS.keys())


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().devias_last_error()
        raise RuntimeError(f'devias_b200.{what} failed (status {rc}): {msg.decode() if msg else "?"}')


def launch_count() -> int:
    return int(lib().devias_launch_count())


def profile_begin(capture_only=False):
    """bracket every launch of the instrumented kernel families with CUDA events from now on; capture_only: only launches made
    inside a stream capture (they become event-record nodes of the graph and are re-recorded by every replay)"""
    check(lib().devias_profile_begin_capture() if capture_only else lib().devias_profile_begin(), 'profile_begin')


def profile_pause():
    check(lib().devias_profile_pause(), 'profile_pause')


def profile_end(kind: int):
    """-> (total device ms, total algorithmic work, launches) of one kernel family since profile_begin()"""
    ms, work, n = ctypes.c_double(), ctypes.c_double(), c_int64()
    check(lib().devias_profile_end(kind, ctypes.byref(ms), ctypes.byref(work), ctypes.byref(n)), 'profile_end')
    return ms.value, work.value, n.value
