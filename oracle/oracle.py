"""TEST INFRASTRUCTURE ONLY — Python face of the CPU oracle (oracle/sucre_oracle.c).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package (sucre_b200/) must never import this module; tests/test_no_oracle_in_product.py greps for it.

Parity is pinned by tests/test_oracle_golden.py against outputs of the unmodified reference
(tests/golden/*.npz, produced by oracle/gen_golden.py).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import torch

_HERE = Path(__file__).resolve().parent
_LIB = None


class OracleView(C.Structure):
    _fields_ = [('K', C.c_float * 9), ('Kinv', C.c_float * 9), ('R', C.c_float * 9), ('t', C.c_float * 3),
                ('Ri', C.c_float * 9), ('ti', C.c_float * 3), ('width', C.c_int32), ('height', C.c_int32)]


def build():
    subprocess.run(['make', '-C', str(_HERE), '--no-print-directory'], check=True, capture_output=True)


def lib():
    global _LIB
    if _LIB is None:
        so = _HERE / '_oracle.so'
        if not so.exists():
            build()
        L = C.CDLL(str(so))
        L.oracle_match_pair.restype = C.c_int64
        L.oracle_match_pair.argtypes = [C.c_void_p, C.POINTER(OracleView), C.c_void_p, C.POINTER(OracleView),
                                        C.c_void_p, C.POINTER(C.c_int64)]
        L.oracle_sample_pair.restype = C.c_int64
        L.oracle_sample_pair.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(OracleView)] \
            + [C.c_void_p] * 8 + [C.c_int64]
        L.oracle_fit.restype = C.c_int
        L.oracle_fit.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                 C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


def view_geom(K, R, t, width: int, height: int, Kinv=None, Ri=None, ti=None) -> OracleView:
    """Per-view constants.  K (3,3), R (3,3), t (3,1) are the reference's Camera.K and cam->world Pose
    (sfm.py:204-208, 219-222); the derived matrices use the reference's own torch expressions
    (K.inverse() sfm.py:92, Pose.inverse() sfm.py:47) unless given (golden fixtures carry them)."""
    K = torch.as_tensor(np.asarray(K), dtype=torch.float32).reshape(3, 3)
    R = torch.as_tensor(np.asarray(R), dtype=torch.float32).reshape(3, 3)
    t = torch.as_tensor(np.asarray(t), dtype=torch.float32).reshape(3, 1)
    Kinv = K.inverse() if Kinv is None else torch.as_tensor(np.asarray(Kinv), dtype=torch.float32).reshape(3, 3)
    Ri = R.T if Ri is None else torch.as_tensor(np.asarray(Ri), dtype=torch.float32).reshape(3, 3)
    ti = (-R.T @ t) if ti is None else torch.as_tensor(np.asarray(ti), dtype=torch.float32).reshape(3, 1)
    g = OracleView()
    for name, val in (('K', K), ('Kinv', Kinv), ('R', R), ('t', t), ('Ri', Ri), ('ti', ti)):
        flat = val.contiguous().flatten().tolist()
        getattr(g, name)[:] = flat
    g.width, g.height = int(width), int(height)
    return g


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def match_pair(depthT: np.ndarray, gT: OracleView, depthS: np.ndarray, gS: OracleView):
    """Dense two-way match.  Returns (idx int32 (H_T,W_T): u2 | v2<<16 or -1, n_matches, n_inbounds)."""
    depthT = np.ascontiguousarray(depthT, dtype=np.uint16)
    depthS = np.ascontiguousarray(depthS, dtype=np.uint16)
    assert depthT.shape == (gT.height, gT.width) and depthS.shape == (gS.height, gS.width)
    idx = np.empty((gT.height, gT.width), dtype=np.int32)
    nin = C.c_int64(0)
    n = lib().oracle_match_pair(_ptr(depthT), C.byref(gT), _ptr(depthS), C.byref(gS), _ptr(idx), C.byref(nin))
    return idx, int(n), int(nin.value)


def sample_pair(idx: np.ndarray, depthS: np.ndarray, rgbS: np.ndarray | None, gS: OracleView) -> dict:
    """Compacted observation arrays of one view, in the reference's layout and dtypes."""
    H, W = idx.shape
    n = int((idx >= 0).sum())
    depthS = np.ascontiguousarray(depthS, dtype=np.uint16)
    out = dict(u1=np.empty(n, np.int16), v1=np.empty(n, np.int16), u2=np.empty(n, np.int16),
               v2=np.empty(n, np.int16), d=np.empty(n, np.float32), cP=np.empty((3, n), np.float32),
               z=np.empty(n, np.float32), I=np.empty((3, n), np.float32))
    rgb_ptr = None
    if rgbS is not None:
        rgbS = np.ascontiguousarray(rgbS, dtype=np.uint8)
        assert rgbS.shape == (gS.height, gS.width, 3)
        rgb_ptr = _ptr(rgbS)
    k = lib().oracle_sample_pair(_ptr(np.ascontiguousarray(idx)), W, H, _ptr(depthS), rgb_ptr, C.byref(gS),
                                 *[_ptr(out[key]) for key in ('u1', 'v1', 'u2', 'v2', 'd', 'cP', 'z', 'I')], n)
    assert k == n
    return out


def gather(depthT: np.ndarray, gT: OracleView, sources: list, min_cover: float = 1e-6, sample: bool = True):
    """Restates Image.match_images + prepare_matches + load_matches (sfm.py:127-138, loader.py:78-118).
    `sources` = [(key, depth u16, rgb u8 or None, OracleView)], already in the order observations are to be
    consumed (the reference iterates kept views sorted by name).  Returns (kept list of (key, obs dict),
    stats dict key -> (n, n_inbounds))."""
    kept, stats = [], {}
    for key, depthS, rgbS, gS in sources:
        idx, n, nin = match_pair(depthT, gT, depthS, gS)
        stats[key] = (n, nin)
        if n / (gT.width * gT.height) > min_cover:  # sfm.py:136, python float division
            kept.append((key, sample_pair(idx, depthS, rgbS, gS) if sample else dict(idx=idx)))
    return kept, stats


def fit(obs_per_view: list[dict], width: int, height: int, *, closed_form: bool = True, num_iter: int = 200,
        lr: float = 0.05, params=None, J0: np.ndarray | None = None):
    """Restates SUCRe + adam (sucre.py:35-82, 124-157).  obs_per_view: dicts with u1, v1, z, I (3,n).
    Returns dict(params (9,), history (num_iter,9), cost (num_iter,), J (H,W,3))."""
    P = width * height
    off = np.zeros(len(obs_per_view) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(o['u1']) for o in obs_per_view])
    pix = np.concatenate([o['v1'].astype(np.int32) * width + o['u1'].astype(np.int32) for o in obs_per_view]) \
        if obs_per_view else np.zeros(0, np.int32)
    z = np.concatenate([o['z'] for o in obs_per_view]).astype(np.float32)
    I = np.ascontiguousarray(np.concatenate([o['I'].T for o in obs_per_view]).astype(np.float32))
    p = np.full(9, 0.1, dtype=np.float32) if params is None else np.array(params, dtype=np.float32).reshape(9)
    if closed_form:
        J = np.empty((height, width, 3), dtype=np.float32)
    else:
        assert J0 is not None
        J = np.ascontiguousarray(J0, dtype=np.float32).copy()
    history = np.zeros((num_iter, 9), dtype=np.float32)
    cost = np.zeros(num_iter, dtype=np.float64)
    rc = lib().oracle_fit(0 if closed_form else 1, len(obs_per_view), _ptr(off), _ptr(pix), _ptr(z), _ptr(I), P,
                          _ptr(p), _ptr(J), num_iter, lr, _ptr(history), _ptr(cost))
    assert rc == 0
    return dict(params=p, history=history, cost=cost, J=J)


def initial_J(rgb_u8: np.ndarray, depth_u16: np.ndarray) -> np.ndarray:
    """sucre.py:47-49: J starts as the target image, NaN where target depth <= 0."""
    J = (rgb_u8.astype(np.float32) / np.float32(255.0)).astype(np.float32)
    J[depth_u16 == 0] = np.nan
    return J
