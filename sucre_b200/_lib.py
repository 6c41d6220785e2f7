"""ctypes binding of libsucre_b200.so (include/sucre_b200.h).  No fallback: if the library is missing or a call
fails, an exception is raised."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / 'libsucre_b200.so'
ABI_VERSION = 10
TILE = 32
REC_Z_U8, REC_Z_F32, REC_P_U8, REC_P_F32 = 0, 1, 2, 3
RECORD_BYTES = {REC_Z_U8: 8, REC_Z_F32: 16, REC_P_U8: 16, REC_P_F32: 32}
FIT_CLOSED_FORM, FIT_PARAM_J = 0, 1

# numpy mirror of `struct sucre_view` (208 bytes)
VIEW_DTYPE = np.dtype([('K', '<f4', 9), ('Kinv', '<f4', 9), ('R', '<f4', 9), ('t', '<f4', 3), ('Ri', '<f4', 9),
                       ('ti', '<f4', 3), ('width', '<i4'), ('height', '<i4'), ('depth', '<u8'), ('rgb', '<u8'),
                       ('rgb_format', '<i4'), ('flags', '<i4'), ('reserved', '<i4', 2)])
assert VIEW_DTYPE.itemsize == 208
RGB_U8, RGB_F32 = 0, 1
VIEW_K_SPARSE, VIEW_KINV_SPARSE = 1, 2


class SucreError(RuntimeError):
    pass


class SucreStore(C.Structure):
    """ctypes mirror of `struct sucre_store` (host struct of device pointers, 48 bytes)."""
    _fields_ = [('cells', C.c_void_p), ('row_off', C.c_void_p), ('n_tiles', C.c_int32), ('record_format', C.c_int32),
                ('pixels', C.c_int64), ('n_rows', C.c_int64), ('pix', C.c_void_p)]


assert C.sizeof(SucreStore) == 48
GROUP_TILES = 32   # SUCRE_GROUP_TILES: tiles whose pixels sucre_gather_permute deals by observation count


class Band(C.Structure):
    """ctypes mirror of `struct sucre_band`: the tiles of a target one call (one rank) covers.  Local tile k is the
    target's tile first_tile + (k // chunk_tiles) * stride_tiles + k % chunk_tiles."""
    _fields_ = [('first_tile', C.c_int32), ('n_tiles', C.c_int32), ('chunk_tiles', C.c_int32), ('stride_tiles', C.c_int32)]

    @staticmethod
    def whole(n_tiles_total: int) -> 'Band':
        return Band(0, n_tiles_total, n_tiles_total, 0)

    @staticmethod
    def contiguous(n_tiles_total: int, rank: int, world: int) -> 'Band':
        """Equal consecutive runs of tiles (balanced to within one tile)."""
        lo = n_tiles_total * rank // world
        n = n_tiles_total * (rank + 1) // world - lo
        return Band(lo, n, max(n, 1), 0)

    @staticmethod
    def cyclic(n_tiles_total: int, rank: int, world: int, chunk: int = 64) -> 'Band':
        """Chunks of `chunk` tiles dealt round-robin: every rank sees every region of the image, so observation counts
        balance to ~1 % where contiguous bands differ by ~10 % (tools/band_balance.py)."""
        period = world * chunk
        full, rem = divmod(n_tiles_total, period)
        n = full * chunk + min(max(rem - rank * chunk, 0), chunk)
        return Band(rank * chunk, n, chunk, period)

    def as_tuple(self) -> tuple:
        return (self.first_tile, self.n_tiles, self.chunk_tiles, self.stride_tiles)

    def tile(self, k: int) -> int:
        """Global tile index of local tile k."""
        return self.first_tile + (k // self.chunk_tiles) * self.stride_tiles + k % self.chunk_tiles

    def tiles(self) -> np.ndarray:
        """Global tile index of every local tile."""
        k = np.arange(self.n_tiles, dtype=np.int64)
        return self.first_tile + (k // self.chunk_tiles) * self.stride_tiles + k % self.chunk_tiles

    def pixels(self, target_pixels: int) -> np.ndarray:
        """Global flat pixel index of every local pixel that exists in the image (local order)."""
        p = (self.tiles()[:, None] * TILE + np.arange(TILE, dtype=np.int64)[None, :]).reshape(-1)
        return p[p < target_pixels]


assert C.sizeof(Band) == 16
MAX_PEERS, PEER_BUFFER_BYTES = 16, 5120


_lib = None

_VP, _I, _I64, _D = C.c_void_p, C.c_int, C.c_int64, C.c_double
_SIGNATURES = {
    'sucre_abi_version': (C.c_int, []),
    'sucre_record_bytes': (C.c_int, [_I]),
    'sucre_last_error': (C.c_char_p, []),
    'sucre_scene_upload': (C.c_int, [_VP, _VP, _I, _VP, _I, _I, _I, _VP, _VP, _VP]),
    'sucre_gather_match': (C.c_int, [_VP, _VP, _I, _VP, _VP, _VP, _VP]),
    'sucre_gather_count': (C.c_int, [_VP, _I, _I, _VP, _VP]),
    'sucre_gather_plan': (C.c_int, [_VP, _I, _I, _VP, _I64, _D, _VP, _VP, _VP, _VP, _VP, _VP]),
    'sucre_gather_permute': (C.c_int, [_VP, _I, _VP, _I64, _VP, _VP, _VP]),
    'sucre_gather_sample': (C.c_int, [_VP, _VP, _I, _VP, _VP, _VP, _VP, _VP, _VP, _I, _VP, _VP, _VP, _VP, _VP]),
    'sucre_band_scatter_J': (C.c_int, [_VP, _VP, _VP, _I64, _VP, _I, _VP]),
    'sucre_fit_workspace_bytes': (C.c_size_t, []),
    'sucre_fit_prepare': (C.c_int, [_VP, _VP, _VP]),
    'sucre_fit_sums': (C.c_int, [_I, _VP, _VP, _VP, _VP, _I64, _I, _D, _VP, _VP, _VP]),
    'sucre_adam_step': (C.c_int, [_VP, _VP, _VP, _I64, _I, _D, _VP, _VP]),
    'sucre_fit': (C.c_int, [_I, _VP, _I64, _VP, _VP, _VP, _VP, _I, _I, _D, _VP, _VP, _VP]),
    'sucre_fit_sharded': (C.c_int, [_I, _VP, _I64, _VP, _VP, _VP, _VP, _I, _I, _D, _VP, _VP, _VP, _I, _I, C.c_uint32, _VP]),
    'sucre_fit_status': (C.c_int, [_VP, _VP, _VP]),
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
    rec['flags'] = (VIEW_K_SPARSE if _pinhole_sparse(rec['K']) else 0) | (VIEW_KINV_SPARSE if _pinhole_sparse(rec['Kinv']) else 0)
    return rec


def _pinhole_sparse(M) -> bool:
    """True iff the 3x3 matrix is exactly [[a,0,b],[0,c,d],[0,0,1]] (decided on the values: the kernels then skip the
    multiplications by 0 and 1, which cannot change a rounded result)."""
    m = np.asarray(M, dtype=np.float32).reshape(9)
    return bool(m[1] == 0 and m[3] == 0 and m[6] == 0 and m[7] == 0 and m[8] == 1 and np.isfinite(m).all())
