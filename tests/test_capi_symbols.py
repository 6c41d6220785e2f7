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


def test_fit_kernel_sass_keeps_its_shape():
    """Guards what the fit kernel's speed rests on, on the built library (cuobjdump, no GPU needed): 1-D TMA bulk copies
    and mbarriers, packed fp32x2 arithmetic, MUFU.EX2, no tensor-core and no local-memory traffic inside the row loop, and
    a row loop of four two-row steps at no more than 93 instructions per step (ptxas drifts to 98 and spills the next
    tile's J when the register budget of the sweep is disturbed)."""
    import re
    import shutil
    import subprocess
    if shutil.which('cuobjdump') is None:
        pytest.skip('cuobjdump not installed')
    out = subprocess.run(['cuobjdump', '-sass', str(_lib.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    m = re.search(r'Function : (\S*fit_kernelILi0ELi0ELb0\S*)', out)
    assert m, 'closed-form u8 fit kernel not found'
    txt = out[m.start():out.find('Function :', m.start() + 10)]
    ins = [(int(a, 16), t.strip()) for a, t in re.findall(r'/\*([0-9a-f]{4})\*/\s+(.*?);', txt)]
    ops = [t.split()[1] if t.startswith('@') else t.split()[0] for _, t in ins]
    assert any(o.startswith('UBLKCP') for o in ops) and any(o.startswith('SYNCS') for o in ops)
    assert not any(o.startswith(('HMMA', 'UTCHMMA', 'UTCQMMA', 'IMMA')) for o in ops)
    addr = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, t) in enumerate(ins):
        b = re.search(r'BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)', t)
        if b and int(b.group(1), 16) < a and int(b.group(1), 16) in addr:
            body = ops[addr[int(b.group(1), 16)]:i + 1]
            if sum(o.startswith('FFMA2') for o in body) >= 30:
                loops.append(body)
    assert loops, 'row loop not found'
    body = min(loops, key=len)
    steps = sum(o.startswith('FFMA2') for o in body) // 36
    assert steps == 4, steps
    assert len(body) <= 93 * steps, len(body)
    assert sum(o.startswith('MUFU.EX2') for o in body) == 12 * steps
    assert not any(o.startswith(('LDL', 'STL', 'LDG', 'STG')) for o in body)
