"""ctypes binding of libsucre_b200.so (include/sucre_b200.h).  No fallback: if the library is missing or a call
fails, an exception is raised."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / 'libsucre_b200.so'
ABI_VERSION = 6
TILE = 32
SEG_HEADER_CELLS = 2
FIT_CLOSED_FORM, FIT_PARAM_J = 0, 1

# numpy mirror of `struct sucre_view` (208 bytes)
VIEW_DTYPE = np.dtype([('K', '<f4', 9), ('Kinv', '<f4', 9), ('R', '<f4', 9), ('t', '<f4', 3), ('Ri', '<f4', 9),
                       ('ti', '<f4', 3), ('width', '<i4'), ('height', '<i4'), ('depth', '<u8'), ('rgb', '<u8'),
                       ('rgb_format', '<i4'), ('reserved', '<i4', 3)])
assert VIEW_DTYPE.itemsize == 208
RGB_U8, RGB_F32 = 0, 1


class SucreError(RuntimeError):
    pass


class SucreStore(C.Structure):
    """ctypes mirror of `struct sucre_store` (host struct of device pointers, 48 bytes)."""
    _fields_ = [('cells', C.c_void_p), ('rec_off', C.c_void_p), ('blk_off', C.c_void_p), ('seg_off', C.c_void_p),
                ('n_tiles', C.c_int32), ('seg_views', C.c_int32), ('pixels', C.c_int64),
                ('record_cells', C.c_int32), ('reserved', C.c_int32)]


assert C.sizeof(SucreStore) == 56
MAX_PEERS, PEER_BUFFER_BYTES = 16, 3072
LIGHT_SEG_VIEWS = 7   # two-cell records: 2 + 2*32*7 = 450 cells per segment at most


_lib = None

_VP, _I, _I64, _D = C.c_void_p, C.c_int, C.c_int64, C.c_double
_SIGNATURES = {
    'sucre_abi_version': (C.c_int, []),
    'sucre_segment_views': (C.c_int, []),
    'sucre_last_error': (C.c_char_p, []),
    'sucre_scene_upload': (C.c_int, [_VP, _VP, _I, _VP, _I, _I, _I, _VP, _VP, _VP]),
    'sucre_gather_match': (C.c_int, [_VP, _VP, _I, _I, _I, _VP, _VP]),
    'sucre_gather_count': (C.c_int, [_VP, _I, _I, _VP, _VP]),
    'sucre_gather_plan': (C.c_int, [_VP, _I, _I, _VP, _I64, _D, _I, _VP, _VP, _VP, _VP, _VP, _VP]),
    'sucre_gather_sample': (C.c_int, [_VP, _VP, _I, _I, _I, _VP, _VP, _VP, _VP, _VP, _I, _I, _VP, _VP, _VP, _VP, _VP]),
    'sucre_fit_workspace_bytes': (C.c_size_t, []),
    'sucre_fit_prepare': (C.c_int, [_VP, _VP, _VP]),
    'sucre_fit_sums': (C.c_int, [_I, _VP, _VP, _VP, _VP, _I64, _I, _D, _VP, _VP, _VP]),
    'sucre_adam_step': (C.c_int, [_VP, _VP, _VP, _I64, _I, _D, _VP, _VP]),
    'sucre_fit': (C.c_int, [_I, _VP, _I64, _VP, _VP, _VP, _VP, _I, _I, _D, _VP, _VP, _VP]),
    'sucre_fit_sharded': (C.c_int, [_I, _VP, _I64, _VP, _VP, _VP, _VP, _I, _I, _D, _VP, _VP, _VP, _I, _I, C.c_uint32, _VP]),
    'sucre_fit_write_J': (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP]),
    'sucre_light_J': (C.c_int, [_VP, _VP, _VP, _VP]),
    'sucre_light_sums': (C.c_int, [_I, _VP, _VP, _VP, _VP, _I64, _I, _D, _VP, _VP, _VP]),
}
EXPORTS = tuple(_SIGNATURES)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise SucreError(f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                             f'or `make -C sucre_b200/csrc`. sucre_b200 has no CPU fallback.')
        L = C.CDLL(str(LIB_PATH))
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        if L.sucre_abi_version() != ABI_VERSION:
            raise SucreError(f'ABI mismatch: library {L.sucre_abi_version()}, python {ABI_VERSION}')
        _lib = L
    return _lib


def seg_views() -> int:
    """Source views per segment of the observation store (SUCRE_SEGMENT_VIEWS of the loaded library)."""
    return lib().sucre_segment_views()


def check(rc: int, what: str):
    if rc != 0:
        raise SucreError(f'{what}: {lib().sucre_last_error().decode()}')


def view_record(K, Kinv, R, t, Ri, ti, width: int, height: int, depth_ptr: int = 0, rgb_ptr: int = 0,
                rgb_format: int = RGB_U8) -> np.ndarray:
    rec = np.zeros((), dtype=VIEW_DTYPE)
    for name, val in (('K', K), ('Kinv', Kinv), ('R', R), ('t', t), ('Ri', Ri), ('ti', ti)):
        rec[name] = np.asarray(val, dtype=np.float32).reshape(-1)
    rec['width'], rec['height'], rec['depth'], rec['rgb'] = width, height, depth_ptr, rgb_ptr
    rec['rgb_format'] = rgb_format
    return rec
