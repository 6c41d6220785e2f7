"""CPU-only: the C-ABI library loads and exports every symbol include/sucre_b200.h declares; the product never
imports the oracle."""
import ctypes
import re
from pathlib import Path

from sucre_b200 import _lib

ROOT = Path(__file__).resolve().parents[1]


def test_header_symbols_are_exported():
    header = (ROOT / 'include' / 'sucre_b200.h').read_text()
    declared = set(re.findall(r'\b(sucre_[A-Za-z_0-9]+)\s*\(', header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    L = _lib.lib()  # binds every symbol; AttributeError if one is missing
    assert L.sucre_abi_version() == _lib.ABI_VERSION
    assert L.sucre_fit_workspace_bytes() > 0
    assert L.sucre_last_error() == b''


def test_view_struct_layout_matches_header():
    header = (ROOT / 'include' / 'sucre_b200.h').read_text()
    assert 'float K[9], Kinv[9], R[9], t[3], Ri[9], ti[3];' in header
    assert _lib.VIEW_DTYPE.itemsize == 208
    assert [_lib.VIEW_DTYPE.fields[n][1] for n in ('K', 'Kinv', 'R', 't', 'Ri', 'ti', 'width', 'height', 'depth', 'rgb')] \
        == [0, 36, 72, 108, 120, 156, 168, 172, 176, 184]
    assert _lib.VIEW_DTYPE.fields["rgb_format"][1] == 192 and _lib.VIEW_DTYPE.fields["flags"][1] == 196


def test_store_struct_and_record_formats_match_header():
    header = (ROOT / 'include' / 'sucre_b200.h').read_text()
    L = _lib.lib()
    for name, fmt in (('Z_U8', _lib.REC_Z_U8), ('Z_F32', _lib.REC_Z_F32), ('P_U8', _lib.REC_P_U8), ('P_F32', _lib.REC_P_F32)):
        assert f'#define SUCRE_REC_{name} {fmt}' in header
        assert L.sucre_record_bytes(fmt) == _lib.RECORD_BYTES[fmt]
    assert L.sucre_record_bytes(17) == 0
    assert ctypes.sizeof(_lib.SucreStore) == 48 and 'sizeof == 48' in header
    assert [getattr(_lib.SucreStore, f).offset for f in ('cells', 'row_off', 'n_tiles', 'record_format', 'pixels', 'n_rows', 'pix')] \
        == [0, 8, 16, 20, 24, 32, 40]
    assert f'#define SUCRE_GROUP_TILES {_lib.GROUP_TILES}' in header


def test_pinhole_flags_come_from_the_values():
    import numpy as np
    K = np.array([[900., 0, 320], [0, 900, 240], [0, 0, 1]], dtype=np.float32)
    Kinv = np.linalg.inv(K).astype(np.float32)
    rec = _lib.view_record(K, Kinv, np.eye(3), np.zeros(3), np.eye(3), np.zeros(3), 640, 480)
    assert rec['flags'] == (_lib.VIEW_K_SPARSE | _lib.VIEW_KINV_SPARSE)
    K2 = K.copy()
    K2[0, 1] = 1e-3   # a skew term: the general path must be used
    assert _lib.view_record(K2, Kinv, np.eye(3), np.zeros(3), np.eye(3), np.zeros(3), 640, 480)['flags'] == _lib.VIEW_KINV_SPARSE


def test_argument_errors_without_gpu():
    L = _lib.lib()
    assert L.sucre_gather_plan(0, 1, 1, 0, 1, 0.0, 0, 0, 0, 0, 0, 0) != 0
    assert b'null' in L.sucre_last_error()
    assert L.sucre_adam_step(0, 0, 0, 1, 1, 0.05, 0, 0) != 0
    assert L.sucre_gather_permute(0, 1, 0, 1, 0, 0, 0) != 0 and b'null' in L.sucre_last_error()
    band = _lib.Band.whole(4)
    import ctypes as C
    buf = (C.c_uint32 * 8)()
    assert L.sucre_gather_permute(buf, 1, C.byref(band), 128, buf, buf, 0) != 0 and b'alias' in L.sucre_last_error()
    assert L.sucre_gather_permute(buf, 1 << 21, C.byref(band), 128, buf, (C.c_uint32 * 8)(), 0) != 0 and b'sizes' in L.sucre_last_error()


def test_product_never_imports_the_oracle():
    for path in (ROOT / 'sucre_b200').rglob('*'):
        if path.suffix in ('.py', '.cu', '.cuh', '.h'):
            assert 'oracle' not in path.read_text().lower(), f'{path} mentions the oracle'
