"""Image decode and the matches container, with the reference's call surface (reference: sucre/loader.py).

The reference spills matches to an HDF5 file and streams them back view-batch by view-batch every iteration
(loader.py:36-130).  Here the matches live in an engine.ObservationStore on the device; MatchesFile keeps the
reference's method names so sucre.restore_image reads the same, and its on-disk form (only written for
--keep-matches, only read to honour the reference's "reuse an existing matches file" rule, sucre.py:185) is a
.npz dump of the store because h5py is not a dependency here.
"""
from __future__ import annotations

from pathlib import Path

import cv2
import numpy as np
import torch
from torch import Tensor

from .engine import ObservationStore


# ---- decode ---------------------------------------------------------------------------------------------------
def _imread(path: Path, flags=None) -> np.ndarray:
    img = cv2.imread(str(path)) if flags is None else cv2.imread(str(path), flags)
    if img is None:
        raise FileNotFoundError(f'cannot read image {path}')
    return img


def load_rgb(rgb_path: Path, width: int, height: int) -> Tensor:
    """(H,W,3) float32 in [0,1], value-for-value what loader.py:156-163 returns (kept for API parity; the hot path
    uploads load_rgb_device_form: the raw u8 image, divided by 255 in-kernel, bit-identical for every u8 code)."""
    rgb = cv2.cvtColor(_imread(rgb_path), cv2.COLOR_BGR2RGB) / 255
    if (rgb.shape[0] != height) or (rgb.shape[1] != width):
        rgb = cv2.resize(rgb, (width, height), interpolation=cv2.INTER_AREA if width < rgb.shape[1] else cv2.INTER_CUBIC)
    return torch.tensor(rgb, dtype=torch.float32)


def load_depth_map(depth_map_path: Path, width: int, height: int) -> Tensor:
    """(H,W) float32 metres, value-for-value what loader.py:166-170 returns."""
    depth_map = _imread(depth_map_path, cv2.IMREAD_UNCHANGED) / 1000
    if (depth_map.shape[0] != height) or (depth_map.shape[1] != width):
        depth_map = cv2.resize(depth_map, (width, height), interpolation=cv2.INTER_NEAREST)
    return torch.tensor(depth_map, dtype=torch.float32)


def load_rgb_device_form(rgb_path: Path, width: int, height: int) -> Tensor:
    """Colour in the form the device-resident scene keeps it: (H,W,3) uint8 exactly as stored when the file has the
    camera's size (the /255 then happens in-kernel, bit-identical), else (H,W,3) float32 resampled exactly like
    loader.py:157-163 (float64 /255, INTER_AREA when shrinking else INTER_CUBIC, then float32)."""
    bgr = _imread(rgb_path)
    if (bgr.shape[0] == height) and (bgr.shape[1] == width):
        return torch.from_numpy(np.ascontiguousarray(bgr[..., ::-1]))
    rgb = cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB) / 255
    rgb = cv2.resize(rgb, (width, height), interpolation=cv2.INTER_AREA if width < rgb.shape[1] else cv2.INTER_CUBIC)
    return torch.tensor(rgb, dtype=torch.float32)


def load_depth_u16(depth_map_path: Path, width: int, height: int) -> Tensor:
    """(H,W) uint16 millimetres.  A size mismatch is resolved with nearest-neighbour like loader.py:168-169;
    nearest resampling commutes with the /1000 scaling, so the u16 grid is preserved."""
    depth = _imread(depth_map_path, cv2.IMREAD_UNCHANGED)
    if depth.ndim != 2:
        raise ValueError(f'{depth_map_path}: expected a single-channel depth map, got shape {depth.shape}')
    if depth.dtype == np.uint8:      # the reference divides whatever integer the file holds by 1000 (loader.py:167)
        depth = depth.astype(np.uint16)
    elif depth.dtype != np.uint16:   # float / 32-bit maps have no exact 16-bit millimetre form
        raise ValueError(f'{depth_map_path}: expected an 8- or 16-bit integer depth map in millimetres, got {depth.dtype}')
    if (depth.shape[0] != height) or (depth.shape[1] != width):
        depth = cv2.resize(depth, (width, height), interpolation=cv2.INTER_NEAREST)
    return torch.from_numpy(np.ascontiguousarray(depth))


# ---- output ---------------------------------------------------------------------------------------------------
class AsyncWriter:
    """A few threads that encode PNGs / dump .pt files while the GPU works on the next target (PNG encoding and
    torch.save release the GIL).  Once the kernels take milliseconds, writing the reference's per-target files
    (sucre.py:116-121, 213-215) is what bounds a multi-target run.  close() waits for every job and re-raises the
    first failure."""

    def __init__(self, num_threads: int | None = None):
        from concurrent.futures import ThreadPoolExecutor
        import os
        num_threads = num_threads or min(8, os.cpu_count() or 1)
        self._pool = ThreadPoolExecutor(max_workers=num_threads, thread_name_prefix="sucre-writer")
        self._jobs = []

    def submit(self, fn, *args):
        self._jobs.append(self._pool.submit(fn, *args))

    def close(self):
        try:
            for job in self._jobs:
                job.result()
        finally:
            self._jobs.clear()
            self._pool.shutdown(wait=True)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


# ---- matches ----------------------------------------------------------------------------------------------------
class MatchesData:
    """What adam() consumes (reference: loader.py:36-53, a list of per-view samples).  Here: a handle on the
    device-resident observation store; len() is the number of observations, as in the reference."""

    def __init__(self, store: ObservationStore, names: list[str]):
        self.store = store
        self.names = names  # image names of store.source_keys

    def __len__(self) -> int:
        return self.store.n_obs

    def iter(self, batch_size: int = 1, device: str = 'cpu'):
        """Reference-shaped view batches (loader.py:43-50) for inspection: yields (u, v, cP, I) with u, v int64 target
        pixel coordinates, I (3, n), and cP (3, n) the observation's camera-frame point when the store keeps it
        (matches computed for the light model).  The default store keeps only the range z = cP.norm(dim=0) — the
        only use the model makes of cP when light_model is off (sucre.py:53) — and then yields cP = (0, 0, z), whose
        norm is z, so SUCRe.forward(u, v, cP) evaluates exactly as on the reference's output.  Not used by the CUDA fit."""
        per_view = list(self.store.to_reference_layout().values())
        for i in range(0, len(per_view), batch_size):
            chunk = per_view[i:i + batch_size]
            if self.store.has_points:
                cP = np.concatenate([c['cP'] for c in chunk], axis=1)
            else:
                z = np.concatenate([c['z'] for c in chunk])
                cP = np.stack([np.zeros_like(z), np.zeros_like(z), z])
            yield (torch.from_numpy(np.concatenate([c['u1'] for c in chunk])).long().to(device),
                   torch.from_numpy(np.concatenate([c['v1'] for c in chunk])).long().to(device),
                   torch.from_numpy(cP).to(device),
                   torch.from_numpy(np.concatenate([c['I'] for c in chunk], axis=1)).to(device))


class MatchesFile:
    def __init__(self, path: Path, colmap_model, overwrite: bool = False):
        """`path` is the reference's `<output>/<image stem>.h5` (sucre.py:179); the cache written next to it is
        `<stem>.matches.npz`."""
        self.path = Path(path)
        self.cache_path = self.path.with_suffix('.matches.npz')
        if overwrite:
            self.path.unlink(missing_ok=True)
            self.cache_path.unlink(missing_ok=True)
        self.colmap_model = colmap_model
        self.store: ObservationStore | None = None
        self.names: list[str] = []

    # -- device-resident side --------------------------------------------------------------------------------
    def set_store(self, store: ObservationStore, names: list[str]):
        self.store, self.names = store, list(names)

    def exists(self) -> bool:
        return self.store is not None or self.cache_path.exists()

    def get_image_list(self) -> list:
        self._ensure_loaded()
        return [self.colmap_model[name] for name, keep in zip(self.names, self.store.view_kept) if keep]

    def prepare_matches(self, num_workers: int = 0):
        """No-op: colour and range are sampled by the fused gather (the reference does it in a second pass over
        the spill file, loader.py:78-87)."""
        self._ensure_loaded()

    def check_integrity(self):
        """The invariants of loader.py:89-101 on the device-resident store: no NaN, colours >= 0, ranges >= 0 (they are
        > 0 by construction: a match requires a positive source depth), consistent offsets.  Evaluated over all record
        slots at once: sentinels are all-zero, and the packed u8 colour word reads as a tiny non-negative float."""
        self._ensure_loaded()
        s = self.store
        if s.n_obs == 0:
            return
        cells = s.cells
        assert not bool(torch.isnan(cells).any()), f'In {self.path}, observations contain NaN(s).'
        if not s.has_points:  # light-model stores also carry camera-frame points, whose x / y are signed
            assert bool((cells >= 0).all()), f'In {self.path}, observations contain invalid value(s).'
        assert int(s.rec_off[-1]) == s.n_obs and int(s.blk_off[-1]) == s.n_blocks and \
            int(s.row_off[-1]) == s.n_rows, f'In {self.path}, corrupt offsets.'

    def load_matches(self, pin_memory: bool = False, device=None) -> MatchesData:
        self._ensure_loaded(device)
        return MatchesData(self.store, self.names)

    def __len__(self) -> int:
        if not self.exists():
            return 0
        self._ensure_loaded()
        return self.store.n_obs

    def __repr__(self) -> str:
        return f'MatchesFile(path={self.path}, {len(self)} observations)'

    # -- disk side (--keep-matches / reuse) -------------------------------------------------------------------
    def save(self):
        s = self.store
        np.savez(self.cache_path, width=s.width, height=s.height, names=np.array(self.names),
                 source_keys=np.array(s.source_keys), view_count=s.view_count, view_kept=s.view_kept,
                 n_obs=s.n_obs, cells=s.cells.cpu().numpy(), rec_off=s.rec_off.cpu().numpy(),
                 blk_off=s.blk_off.cpu().numpy(), row_off=s.row_off.cpu().numpy(), blk_mask=s.blk_mask.cpu().numpy(),
                 blk_view=s.blk_view.cpu().numpy(),
                 cell_src=np.zeros(0, np.int32) if s.cell_src is None else s.cell_src.cpu().numpy(),
                 record_format=s.record_format, pix=np.zeros(0, np.int32) if s.pix is None else s.pix.cpu().numpy())
        self.path.touch()  # the reference's file name marks "matches exist" (sucre.py:185)

    def unlink(self):
        self.path.unlink(missing_ok=True)
        self.cache_path.unlink(missing_ok=True)

    def _ensure_loaded(self, device=None):
        if self.store is not None:
            return
        if not self.cache_path.exists():
            raise FileNotFoundError(f'no matches computed for {self.path}')
        z = np.load(self.cache_path)
        dev = torch.device('cuda' if device is None else device)
        t = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
        self.names = z['names'].tolist()
        self.store = ObservationStore(
            width=int(z['width']), height=int(z['height']), source_keys=tuple(z['source_keys'].tolist()),
            view_count=z['view_count'], view_kept=z['view_kept'], n_obs=int(z['n_obs']),
            n_blocks=int(z['blk_mask'].shape[0]), n_rows=int(z['row_off'][-1]), cells=t(z['cells']),
            rec_off=t(z['rec_off']), blk_off=t(z['blk_off']), row_off=t(z['row_off']), blk_mask=t(z['blk_mask']),
            blk_view=t(z['blk_view']), cell_src=t(z['cell_src']) if z['cell_src'].size else None,
            record_format=int(z['record_format']), pix=t(z['pix']) if 'pix' in z.files and z['pix'].size else None)
