"""Device-resident scene, observation store and the thin wrappers that enqueue the CUDA hot path.

PyTorch is used for device memory, streams and copies only; every computation on the hot path is a kernel of
libsucre_b200.so reached through ctypes (sucre_b200/_lib.py).  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib
from ._lib import TILE


# ------------------------------------------------------------------------------------------------------------
@dataclass
class ViewGeom:
    """Host-side constants of one view, bit-identical to what the reference multiplies by."""
    K: torch.Tensor      # (3,3) Camera.K                       sfm.py:204-208
    Kinv: torch.Tensor   # (3,3) K.inverse()                    sfm.py:92
    R: torch.Tensor      # (3,3) cam->world rotation            sfm.py:219-222
    t: torch.Tensor      # (3,1) cam->world translation
    Ri: torch.Tensor     # (3,3) Pose.inverse().R = R.T         sfm.py:47
    ti: torch.Tensor     # (3,1) Pose.inverse().t = -R.T @ t    sfm.py:47
    width: int
    height: int
    _rec: np.ndarray | None = field(default=None, repr=False, compare=False)   # sucre_view record without pointers
    _f64: tuple | None = field(default=None, repr=False, compare=False)        # (Ri, ti, K) as float64 numpy

    @staticmethod
    def from_pose(K, R, t, width: int, height: int) -> 'ViewGeom':
        """Derives Kinv, Ri, ti on the host with the reference's own torch expressions."""
        K = torch.as_tensor(K, dtype=torch.float32).reshape(3, 3).cpu()
        R = torch.as_tensor(R, dtype=torch.float32).reshape(3, 3).cpu()
        t = torch.as_tensor(t, dtype=torch.float32).reshape(3, 1).cpu()
        return ViewGeom(K=K, Kinv=K.inverse(), R=R, t=t, Ri=R.T, ti=-R.T @ t, width=int(width), height=int(height))

    def record(self, depth_ptr: int = 0, rgb_ptr: int = 0, rgb_format: int = _lib.RGB_U8) -> np.ndarray:
        """The view's `sucre_view` record (the constants are converted once and cached; a ViewGeom is immutable
        once it has been handed to a scene)."""
        rec = self.constants().copy()
        rec['depth'], rec['rgb'], rec['rgb_format'] = depth_ptr, rgb_ptr, rgb_format
        return rec

    def constants(self) -> np.ndarray:
        """The `sucre_view` record with null pointers (cached; do not modify)."""
        if self._rec is None:
            c = lambda x: x.contiguous().numpy()  # noqa: E731
            self._rec = _lib.view_record(c(self.K), c(self.Kinv), c(self.R), c(self.t), c(self.Ri), c(self.ti),
                                         self.width, self.height)
        return self._rec

    def projection_f64(self) -> tuple:
        """(Ri, ti, K) as float64 numpy arrays, for the host-side frustum tests (cached)."""
        if self._f64 is None:
            self._f64 = tuple(x.double().numpy() for x in (self.Ri, self.ti, self.K))
        return self._f64


def projection_stacks(geoms) -> tuple:
    """Stacked float64 constants of a list of views for DeviceScene.footprints / possibly_overlapping:
    (Ri (V,3,3), ti (V,3,1), K (V,3,3), widths (V,), heights (V,))."""
    if len(geoms) == 0:
        z = np.zeros((0, 3, 3))
        return z, np.zeros((0, 3, 1)), z, np.zeros(0, np.int64), np.zeros(0, np.int64)
    f = [g.projection_f64() for g in geoms]
    return (np.stack([x[0] for x in f]), np.stack([x[1] for x in f]), np.stack([x[2] for x in f]),
            np.array([g.width for g in geoms], dtype=np.int64), np.array([g.height for g in geoms], dtype=np.int64))


class DeviceScene:
    """Views resident in HBM: u16 millimetre depth + u8 RGB per view and their `sucre_view` records.

    Replaces the reference's per-(target, view) PNG re-decode (sfm.py:130-133, loader.py:156-170) and the
    float32 images it keeps: 5 bytes per pixel instead of 16, converted in-kernel."""

    def __init__(self, device: str | torch.device = 'cuda'):
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise _lib.SucreError(f'sucre_b200 runs on CUDA devices only (got {device!r}); there is no CPU path')
        self.geom: dict = {}
        self.depth: dict = {}
        self.rgb: dict = {}
        self._tables: dict = {}
        self._ranges: dict = {}
        self._cull_tables: dict = {}

    def __contains__(self, key):
        return key in self.geom

    def __len__(self):
        return len(self.geom)

    def add_view(self, key, geom: ViewGeom, depth_u16: torch.Tensor, rgb_u8: torch.Tensor | None):
        """depth_u16: (H,W) uint16 (or int16 bit pattern); rgb_u8: (H,W,3) uint8, or float32 in [0,1] for images the
        host resampled (--image-scale); host (ideally pinned) or device."""
        assert depth_u16.shape == (geom.height, geom.width), (depth_u16.shape, geom.height, geom.width)
        assert depth_u16.element_size() == 2
        self.depth[key] = depth_u16.to(self.device, non_blocking=True).contiguous()
        if rgb_u8 is not None:
            assert rgb_u8.shape == (geom.height, geom.width, 3) and rgb_u8.dtype in (torch.uint8, torch.float32)
            self.rgb[key] = rgb_u8.to(self.device, non_blocking=True).contiguous()
        self.geom[key] = geom
        self._tables.clear()
        self._cull_tables.clear()
        self._ranges.pop(key, None)

    def remove_views(self, keys):
        """Forgets views (their device memory is released once no store / table refers to it)."""
        for k in keys:
            self.geom.pop(k, None)
            self.depth.pop(k, None)
            self.rgb.pop(k, None)
            self._ranges.pop(k, None)
        self._tables.clear()
        self._cull_tables.clear()

    def view_bytes(self, key) -> int:
        d, c = self.depth[key], self.rgb.get(key)
        return d.numel() * d.element_size() + (0 if c is None else c.numel() * c.element_size())

    def add_views(self, keys, geoms, depth_u16: torch.Tensor, rgb_u8: torch.Tensor):
        """Bulk form: stacked (V,H,W) / (V,H,W,3) tensors moved with one copy each."""
        d = depth_u16.to(self.device, non_blocking=True)
        c = rgb_u8.to(self.device, non_blocking=True)
        for i, (k, g) in enumerate(zip(keys, geoms)):
            self.geom[k], self.depth[k], self.rgb[k] = g, d[i], c[i]
            self._ranges.pop(k, None)
        self._tables.clear()
        self._cull_tables.clear()

    def allocate_views(self, keys, geoms, depth_dtype=torch.uint16) -> tuple[torch.Tensor, torch.Tensor]:
        """Registers equally sized views whose pixels are still to come (upload_rects): returns the stacked device
        planes (n,H,W) depth — zero = invalid, sfm.py:96 — and (n,H,W,3) colour, which scene.depth[k] / scene.rgb[k]
        alias.  Colour is left uninitialised: the gather reads it only where the source depth is valid."""
        n, W, H = len(keys), geoms[0].width, geoms[0].height
        assert all(g.width == W and g.height == H for g in geoms)
        d = torch.zeros((n, H, W), dtype=torch.int16, device=self.device).view(depth_dtype)
        c = torch.empty((n, H, W, 3), dtype=torch.uint8, device=self.device)
        for i, (k, g) in enumerate(zip(keys, geoms)):
            self.geom[k], self.depth[k], self.rgb[k] = g, d[i], c[i]
            self._ranges.pop(k, None)
        self._tables.clear()
        self._cull_tables.clear()
        return d, c

    def upload_rects(self, planes, depth_u16: torch.Tensor, rgb_u8: torch.Tensor, src_index, rects) -> int:
        """Footprint upload into planes = (d, c) of allocate_views (or slices of them): for plane i only the rectangle
        rects[i] = (x0, y0, x1, y1) of host view src_index[i] of the stacked (N,H,W) / (N,H,W,3) host tensors (pinned
        for asynchronous copies) is copied (sucre_scene_upload, cudaMemcpy2DAsync on the current stream).  With
        rectangles from `footprints` the gather reads nothing outside them, so its results are those of a full
        upload.  Returns the bytes copied."""
        d, c = planes
        n, H, W = (int(x) for x in d.shape)
        assert depth_u16.is_contiguous() and rgb_u8.is_contiguous() and depth_u16.element_size() == 2
        assert rgb_u8.dtype == torch.uint8 and not depth_u16.is_cuda and not rgb_u8.is_cuda
        assert tuple(depth_u16.shape[1:]) == (H, W) and tuple(rgb_u8.shape[1:]) == (H, W, 3) and c.shape[0] == n
        assert d.is_contiguous() and c.is_contiguous()
        idx = np.ascontiguousarray(src_index, dtype=np.int32).reshape(n)
        rc = np.ascontiguousarray(rects, dtype=np.int32).reshape(n, 4)
        assert idx.size == 0 or (idx.min() >= 0 and idx.max() < depth_u16.shape[0] == rgb_u8.shape[0])
        copied, total = C.c_int64(0), 0
        with torch.cuda.device(self.device):
            for dst, src, px in ((d, depth_u16, 2), (c, rgb_u8, 3)):
                _lib.check(_lib.lib().sucre_scene_upload(dst.data_ptr(), src.data_ptr(), n, idx.ctypes.data, W, H, px,
                                                         rc.ctypes.data, C.byref(copied), _stream(self.device)),
                           'sucre_scene_upload')
                total += copied.value
        return total

    def depth_range(self, key) -> tuple[float, float]:
        """(smallest non-zero depth, largest depth) of a view in metres (cached; one small device reduction)."""
        if key not in self._ranges:
            d = self.depth[key].view(torch.int16).to(torch.int32) & 0xffff
            lo = torch.where(d > 0, d, torch.full_like(d, 1 << 16)).min()
            lo, hi = (int(x) for x in torch.stack([lo, d.max()]).cpu())
            self._ranges[key] = (lo / 1000.0, hi / 1000.0)
        return self._ranges[key]

    @staticmethod
    def footprints(target_geom: 'ViewGeom', depth_range: tuple[float, float], source_geoms, margin: int = 2,
                   rows_only: bool = False, stacks: tuple | None = None, band_rows: tuple[int, int] | None = None) -> np.ndarray:
        """Conservative footprint of a target in each source view (float64, host): (V,4) int32 rectangles
        (x0, y0, x1, y1), half-open, that contain every source pixel a valid target pixel can land on, hence every
        source pixel the gather reads (the backward leg and the sampling only touch landing pixels, sfm.py:124, 137).
        Argument of `possibly_overlapping`: the target's back-projected pixels lie in the convex slab spanned by its
        four image-corner rays at its smallest and largest depth; when the eight slab corners are in front of the
        source camera the slab projects inside the convex hull of their projections, so their bounding box (+ `margin`
        pixels for the fp32 arithmetic of the kernels) bounds every landing pixel.  Otherwise the whole image is
        returned.  An empty rectangle (x1 <= x0) means no target pixel can land in the view.
        depth_range = (smallest non-zero, largest) target depth in metres; rows_only widens every rectangle to whole
        image rows (one contiguous block per plane); stacks = projection_stacks(source_geoms) if the caller keeps it;
        band_rows = (v0, v1): bound only the target's pixel rows [v0, v1) (one rank's band of a sharded target)."""
        g = target_geom
        dmin, dmax = depth_range
        Ri, ti, K, Ws, Hs = projection_stacks(source_geoms) if stacks is None else stacks
        V = len(Ws)
        full = np.stack([np.zeros(V, np.int64), np.zeros(V, np.int64), Ws, Hs], axis=1)
        if V == 0 or dmax <= 0 or dmin > dmax:
            return full.astype(np.int32)
        Kinv, R, t = (x.double().numpy() for x in (g.Kinv, g.R, g.t))
        v0, v1 = (0, g.height) if band_rows is None else (max(0, int(band_rows[0])), min(g.height, int(band_rows[1])))
        uv1 = np.array([[0, v0, 1], [g.width, v0, 1], [0, v1, 1], [g.width, v1, 1]], dtype=np.float64).T
        rays = Kinv @ uv1
        slab = np.concatenate([rays * (dmin * 0.999), rays * (dmax * 1.001)], axis=1)   # (3,8) camera frame
        world = R @ slab + t
        c = Ri @ world + ti                                                   # (V,3,8)
        front = (c[:, 2] > 1e-6).all(axis=1)
        p = K @ c
        with np.errstate(divide='ignore', invalid='ignore'):
            x, y = p[:, 0] / p[:, 2], p[:, 1] / p[:, 2]
        ok = front & np.isfinite(x).all(axis=1) & np.isfinite(y).all(axis=1)
        x, y = np.where(ok[:, None], x, 0.0), np.where(ok[:, None], y, 0.0)
        big = float(1 << 30)
        lo = lambda a: np.floor(np.clip(a.min(axis=1), -big, big)).astype(np.int64) - margin           # noqa: E731
        hi = lambda a: np.ceil(np.clip(a.max(axis=1), -big, big)).astype(np.int64) + margin + 1        # noqa: E731
        rect = np.stack([np.clip(lo(x), 0, Ws), np.clip(lo(y), 0, Hs), np.clip(hi(x), 0, Ws), np.clip(hi(y), 0, Hs)], axis=1)
        empty = (rect[:, 2] <= rect[:, 0]) | (rect[:, 3] <= rect[:, 1])
        rect[empty] = 0
        if rows_only:
            rect[~empty, 0], rect[~empty, 2] = 0, Ws[~empty]
        return np.where(ok[:, None], rect, full).astype(np.int32)

    def possibly_overlapping(self, target_key, source_keys, source_geoms=None) -> np.ndarray:
        """Conservative frustum pre-test (float64, host): False only for source views in which NO target pixel can land.
        The target's back-projected pixels all lie in the convex slab spanned by its four image corners at its
        smallest and largest depth; if the eight slab corners are in front of a source camera and their projections
        all fall off the same side of its image (2-pixel margin), no forward projection is in bounds, so the view
        has zero matches whatever its depth map says.  Such views are skipped by the gather and reported in
        ObservationStore.stats['views_culled'].  source_geoms (optional, one ViewGeom per key): lets the test run
        before the source views are decoded / uploaded — only the target has to be resident."""
        g = self.geom[target_key]
        dmin, dmax = self.depth_range(target_key)
        keep = np.ones(len(source_keys), dtype=bool)
        if dmax <= 0 or dmin > dmax:
            return keep
        Kinv, R, t = (x.double().numpy() for x in (g.Kinv, g.R, g.t))
        uv1 = np.array([[0, 0, 1], [g.width, 0, 1], [0, g.height, 1], [g.width, g.height, 1]], dtype=np.float64).T
        rays = Kinv @ uv1                                                     # (3,4)
        slab = np.concatenate([rays * (dmin * 0.999), rays * (dmax * 1.001)], axis=1)   # (3,8) camera frame
        world = R @ slab + t
        keys = tuple(source_keys)
        if keys not in self._cull_tables:  # stacked float64 constants of the listed views, built once
            self._cull_tables[keys] = projection_stacks([self.geom[k] for k in keys] if source_geoms is None else list(source_geoms))
        Ri, ti, K, Ws, Hs = self._cull_tables[keys]
        c = Ri @ world + ti                                                   # (V,3,8)
        front = (c[:, 2] > 1e-6).all(axis=1)
        p = K @ c
        with np.errstate(divide='ignore', invalid='ignore'):
            x, y = p[:, 0] / p[:, 2], p[:, 1] / p[:, 2]
        outside = (x.max(1) < -2) | (x.min(1) > Ws + 1) | (y.max(1) < -2) | (y.min(1) > Hs + 1)
        return ~(front & outside)

    def record(self, key) -> np.ndarray:
        rgb = self.rgb.get(key)
        fmt = _lib.RGB_F32 if rgb is not None and rgb.dtype == torch.float32 else _lib.RGB_U8
        return self.geom[key].record(self.depth[key].data_ptr(), 0 if rgb is None else rgb.data_ptr(), fmt)

    def rgb_float(self, key) -> torch.Tensor:
        """(H,W,3) float32 colour of a view as the reference's load_rgb returns it (loader.py:156-163)."""
        rgb = self.rgb[key]
        return rgb.clone() if rgb.dtype == torch.float32 else rgb.to(torch.float32) / 255.0

    def table(self, keys) -> torch.Tensor:
        """Device array of `sucre_view` for `keys` (cached)."""
        keys = tuple(keys)
        if keys not in self._tables:
            host = np.empty(len(keys), dtype=_lib.VIEW_DTYPE)
            for i, k in enumerate(keys):
                host[i] = self.geom[k].constants()
            rgbs = [self.rgb.get(k) for k in keys]
            host['depth'] = [self.depth[k].data_ptr() for k in keys]
            host['rgb'] = [0 if c is None else c.data_ptr() for c in rgbs]
            host['rgb_format'] = [_lib.RGB_F32 if c is not None and c.dtype == torch.float32 else _lib.RGB_U8 for c in rgbs]
            self._tables[keys] = torch.from_numpy(host.view(np.uint8).reshape(len(keys), -1)).to(self.device)
        return self._tables[keys]


# ------------------------------------------------------------------------------------------------------------
@dataclass
class ObservationStore:
    """Tile-major ELL observation rows (layout: include/sucre_b200.h).  Replaces the reference's HDF5 spill file +
    MatchesData (loader.py:36-130)."""
    width: int
    height: int
    source_keys: tuple
    view_count: np.ndarray        # (V,) int64 matches per listed view over the WHOLE image (host)
    view_kept: np.ndarray         # (V,) bool  min_cover decision (host)
    n_obs: int                    # records in this store
    n_blocks: int
    n_rows: int
    cells: torch.Tensor           # (max(n_rows, 1), 32, record bytes / 4) f32: ELL rows of records (bit patterns; see `records`)
    row_off: torch.Tensor         # (n_tiles+1,) int64 rows before tile k
    blk_off: torch.Tensor         # (n_tiles+1,) int64 blocks before tile k
    rec_off: torch.Tensor         # (n_tiles+1,) int64 records before tile k
    blk_mask: torch.Tensor        # (n_blocks,) int32 (bit pattern of the uint32 lane mask)
    blk_view: torch.Tensor        # (n_blocks,) int32 index into source_keys
    cell_src: torch.Tensor | None  # (n_rows*32,) int32 u2 | v2 << 16 per record slot, -1 at sentinels
    workspace: torch.Tensor | None = None  # fit scratch, prepared on first use
    band: _lib.Band | None = None  # the tiles of the target this store covers (multi-GPU pixel sharding); None = all
    record_format: int = _lib.REC_Z_U8
    stats: dict = field(default_factory=dict)
    pix: torch.Tensor | None = None  # (n_tiles*32,) int32: flat index in the target image of the pixel in every slot, -1 = none
                                     # (sucre_gather_permute); None: slot 32k+i is pixel 32*band.tile(k)+i

    def __post_init__(self):
        if self.band is None:
            self.band = _lib.Band.whole((self.width * self.height + TILE - 1) // TILE)

    @property
    def n_tiles(self) -> int:
        return self.band.n_tiles

    @property
    def is_band(self) -> bool:
        return self.band.as_tuple() != _lib.Band.whole((self.width * self.height + TILE - 1) // TILE).as_tuple()

    @property
    def local_pixels(self) -> int:
        """Target pixels covered by this store, in local order; only the image's last tile can be partial, and it is
        the band's last local tile if the band owns it."""
        P = self.width * self.height
        last_global = self.band.tile(self.band.n_tiles - 1) if self.band.n_tiles else 0
        return self.band.n_tiles * TILE - max(0, (last_global + 1) * TILE - P)

    @property
    def n_slots(self) -> int:
        return self.band.n_tiles * TILE

    def global_pixels(self) -> torch.Tensor:
        """Flat index in the image of the pixel behind every row of a band's J (device int64; -1: a slot without a pixel,
        only in permuted stores)."""
        if self.pix is not None:
            return self.pix.to(torch.int64)
        return torch.from_numpy(self.band.pixels(self.width * self.height)).to(self.cells.device)

    @property
    def J_shape(self) -> tuple:
        """Shape of the J arrays callers see: the image for a whole-image store; for a band, one row per slot of a
        permuted store (see global_pixels) or per local pixel."""
        if not self.is_band:
            return (self.height, self.width, 3)
        return (self.n_slots if self.pix is not None else self.local_pixels, 3)

    @property
    def slot_J_shape(self) -> tuple:
        """Shape of the slot-ordered J arrays the Adam loop works on (include/sucre_b200.h, sucre_store)."""
        return (self.n_slots, 3) if self.pix is not None else self.J_shape

    @property
    def needs_slot_copy(self) -> bool:
        """A caller-visible J of this store (image order) is not what the Adam loop addresses (slot order)."""
        return self.pix is not None and not self.is_band

    def to_slots(self, x: torch.Tensor) -> torch.Tensor:
        """(pixels..., c) image-ordered values -> (n_slots, c) slot order (slots without a pixel get pixel 0's values; the
        kernels never touch them)."""
        return x.reshape(-1, x.shape[-1])[self.pix.to(torch.int64).clamp_(min=0)].contiguous()

    def from_slots(self, slots: torch.Tensor, out: torch.Tensor):
        """Scatter (n_slots, c) slot-ordered values back into the image-ordered `out` (in place)."""
        px = self.pix.to(torch.int64)
        ok = px >= 0
        out.reshape(-1, out.shape[-1])[px[ok]] = slots[ok]

    @property
    def kept_keys(self) -> list:
        return [k for k, keep in zip(self.source_keys, self.view_kept) if keep]

    @property
    def has_points(self) -> bool:
        return self.record_format in (_lib.REC_P_U8, _lib.REC_P_F32)

    @property
    def record_bytes(self) -> int:
        return _lib.RECORD_BYTES[self.record_format]

    @property
    def stream_bytes(self) -> int:
        """Bytes one sweep of the fit reads: every row once."""
        return self.n_rows * TILE * self.record_bytes

    @property
    def fill(self) -> float:
        """Fraction of the record slots that hold an observation (the rest are sentinels below shorter columns)."""
        return self.n_obs / max(1, self.n_rows * TILE)

    def __len__(self) -> int:
        return self.n_obs

    def c_struct(self) -> _lib.SucreStore:
        """Pixel-ordered arrays (sucre_fit_write_J's output, the light model's J) follow J_shape: image order through pix
        for a whole-image store, slot order for a band."""
        if self.pix is None:
            return _lib.SucreStore(self.cells.data_ptr(), self.row_off.data_ptr(), self.n_tiles, self.record_format,
                                   self.local_pixels, self.n_rows, None)
        if self.is_band:
            return _lib.SucreStore(self.cells.data_ptr(), self.row_off.data_ptr(), self.n_tiles, self.record_format,
                                   self.n_slots, self.n_rows, None)
        return _lib.SucreStore(self.cells.data_ptr(), self.row_off.data_ptr(), self.n_tiles, self.record_format,
                               self.width * self.height, self.n_rows, self.pix.data_ptr())

    def record_index(self) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """(slot, pixel, view) of every record (device int64 tensors, block-major order); slot = row * 32 + lane.
        Decoded from the block list: the j-th record of lane i in a tile is the tile's j-th block whose mask has bit
        i.  For export / parity checks, not used by the kernels."""
        dev = self.cells.device
        nblk_tile = self.blk_off[1:] - self.blk_off[:-1]
        blk_tile = torch.repeat_interleave(torch.arange(self.n_tiles, device=dev), nblk_tile)
        lanes = torch.arange(32, device=dev, dtype=torch.int64)
        bits = (self.blk_mask.to(torch.int64)[:, None] >> lanes[None, :]) & 1      # (n_blocks, 32)
        cs = torch.cumsum(bits, dim=0)
        cs_pad = torch.cat([torch.zeros((1, 32), dtype=torch.int64, device=dev), cs])
        j = cs - bits - cs_pad[self.blk_off[blk_tile]]                            # rank within the lane's column
        slot = (self.row_off[blk_tile][:, None] + j) * TILE + lanes[None, :]
        b, lane = torch.nonzero(bits, as_tuple=True)
        if self.pix is not None:
            pixel = self.pix.to(torch.int64)[blk_tile[b] * TILE + lane]
        else:
            pixel = torch.from_numpy(self.band.tiles()).to(dev)[blk_tile[b]] * TILE + lane
        return slot[b, lane], pixel, self.blk_view.to(torch.int64)[b]

    def _flat(self) -> torch.Tensor:
        return self.cells.reshape(-1, self.cells.shape[-1])

    def colours(self, slot: torch.Tensor) -> np.ndarray:
        """(n, 3) float32 I of the records at `slot`, formed like the reference: f32(f64(u8) / 255) (loader.py:157)."""
        flat = self._flat()
        if self.record_format in (_lib.REC_Z_F32, _lib.REC_P_F32):
            cols = slice(1, 4) if self.record_format == _lib.REC_Z_F32 else slice(4, 7)
            return flat[slot][:, cols].cpu().numpy()
        word = flat[slot][:, 1 if self.record_format == _lib.REC_Z_U8 else 3].contiguous().view(torch.int32).cpu().numpy()
        rgb = np.stack([word & 0xff, (word >> 8) & 0xff, (word >> 16) & 0xff], axis=1)
        return (rgb.astype(np.float64) / 255).astype(np.float32)

    def ranges(self, slot: torch.Tensor) -> torch.Tensor:
        """(n,) z = ||cP|| of the records at `slot`."""
        flat = self._flat()
        if self.record_format in (_lib.REC_Z_U8, _lib.REC_Z_F32):
            return flat[slot][:, 0]
        if self.record_format == _lib.REC_P_F32:
            return flat[slot][:, 3]
        c = flat[slot][:, :3]   # sequential squares like cP.norm(dim=0) on the reference's path
        return torch.sqrt(c[:, 0] * c[:, 0] + c[:, 1] * c[:, 1] + c[:, 2] * c[:, 2])

    def records(self) -> torch.Tensor:
        """(n_obs, 4) {z, I_r, I_g, I_b} of every record (device f32, block-major order like record_index)."""
        slot = self.record_index()[0]
        I = torch.from_numpy(self.colours(slot)).to(self.cells.device)
        return torch.cat([self.ranges(slot)[:, None], I], dim=1)

    def to_reference_layout(self) -> dict:
        """Per kept view (in source_keys order) the arrays the reference's MatchesFile/MatchesData hold
        (loader.py:68-76, 103-118): u1, v1, u2, v2 int16, z f32, I (3,n) f32 (cP (3,n) for stores with points) — rows
        ordered row-major over the target like torch.where (sfm.py:96).  Host numpy; for parity tests and --keep-matches."""
        slot, pixel, view = self.record_index()
        out = {}
        flat = self._flat()
        for vi, key in enumerate(self.source_keys):
            if not self.view_kept[vi]:
                continue
            sel = (view == vi).nonzero(as_tuple=True)[0]
            sel = sel[torch.argsort(pixel[sel])]
            p, sl = pixel[sel], slot[sel]
            entry = dict(u1=(p % self.width).to(torch.int16).cpu().numpy(),
                         v1=(p // self.width).to(torch.int16).cpu().numpy(),
                         z=self.ranges(sl).cpu().numpy().copy(), I=np.ascontiguousarray(self.colours(sl).T))
            if self.has_points:
                entry['cP'] = np.ascontiguousarray(flat[sl][:, :3].cpu().numpy().T)
            if self.cell_src is not None:
                src = self.cell_src[sl].cpu().numpy().view(np.uint32)
                entry['u2'] = (src & 0xffff).astype(np.int16)
                entry['v2'] = (src >> 16).astype(np.int16)
            out[key] = entry
        return out


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def gather(scene: DeviceScene, target_key, source_keys, min_cover: float = 1e-6, keep_src: bool = False,
           target_record: np.ndarray | None = None, band: _lib.Band | None = None,
           reduce_counts=None, with_points: bool = False, cull_views: bool | None = None,
           keep_mask: np.ndarray | None = None, permute: bool | None = None) -> ObservationStore:
    """Stage 1 (see _gather_listed) behind a conservative view-level frustum pre-test: source views in which no target
    pixel can land (DeviceScene.possibly_overlapping) are not handed to the kernels at all — on a 1000-view survey a
    target overlaps a few dozen views.  Results are identical with or without it: a culled view has zero matches
    and is reported as not kept; the number of culled views is in stats['views_culled'].  cull_views=None applies
    the test to pairing lists of 128 views or more (below that its host cost exceeds what it can save).
    keep_mask: the result of a pre-test the caller already ran (views marked False need not even be in the scene)."""
    source_keys = tuple(source_keys)
    if cull_views is None:
        cull_views = len(source_keys) >= 128
    kw = dict(min_cover=min_cover, keep_src=keep_src, target_record=target_record, band=band,
              reduce_counts=reduce_counts, with_points=with_points, permute=permute)
    if keep_mask is None and (not cull_views or len(source_keys) < 2 or target_record is not None
                              or target_key not in scene.geom):
        return _gather_listed(scene, target_key, source_keys, **kw)
    keep = scene.possibly_overlapping(target_key, source_keys) if keep_mask is None else np.array(keep_mask, dtype=bool)
    if keep.all():
        return _gather_listed(scene, target_key, source_keys, **kw)
    if not keep.any():
        keep[0] = True  # the kernels need a non-empty list; this view yields zero matches
    idx = np.flatnonzero(keep)
    store = _gather_listed(scene, target_key, [source_keys[i] for i in idx], **kw)
    view_count = np.zeros(len(source_keys), dtype=store.view_count.dtype)
    view_kept = np.zeros(len(source_keys), dtype=bool)
    view_count[idx], view_kept[idx] = store.view_count, store.view_kept
    store.blk_view = torch.from_numpy(idx.astype(np.int32)).to(store.blk_view.device)[store.blk_view.long()]
    store.source_keys, store.view_count, store.view_kept = source_keys, view_count, view_kept
    store.stats['views_culled'] = int(len(source_keys) - len(idx))
    return store


def _gather_listed(scene: DeviceScene, target_key, source_keys, min_cover: float = 1e-6, keep_src: bool = False,
                   target_record: np.ndarray | None = None, band: _lib.Band | None = None,
                   reduce_counts=None, with_points: bool = False, permute: bool | None = None) -> ObservationStore:
    """Stage 1 on the device: match -> count -> permute -> plan -> (one 40-byte D2H to size the store) -> sample.
    Replaces Image.match_images + MatchesFile.prepare_matches + load_matches
    (sfm.py:127-138, loader.py:78-87, 103-118).

    band (a _lib.Band) restricts the call to a band of the target's tiles (multi-GPU pixel sharding);
    reduce_counts(view_count) then sums the per-view match counts over all bands in place (an all-reduce), because
    min_cover is a whole-image criterion (sfm.py:136).
    with_points: keep the camera-frame point cP of every observation, which the light model needs (sucre.py:57); the
    default store keeps only its norm.  Colour stays u8 in the records unless a listed view carries float colour
    (--image-scale).
    permute (default on; SUCRE_PERMUTE=0 turns it off): within every group of 32 tiles the pixels are dealt to the slots
    by decreasing match count (sucre_gather_permute), which removes most sentinels from the ELL rows; the store then
    carries the slot -> pixel map `pix`."""
    L = _lib.lib()
    dev = scene.device
    source_keys = tuple(source_keys)
    V = len(source_keys)
    if V == 0:
        raise _lib.SucreError('gather: empty pairing list')
    trec = scene.record(target_key) if target_record is None else target_record
    W, H = int(trec['width']), int(trec['height'])
    P = W * H
    band = _lib.Band.whole((P + TILE - 1) // TILE) if band is None else band
    n_tiles = band.n_tiles
    table = scene.table(source_keys)
    f32_colour = any(scene.rgb.get(k) is not None and scene.rgb[k].dtype == torch.float32 for k in source_keys)
    fmt = (_lib.REC_P_F32 if f32_colour else _lib.REC_P_U8) if with_points else (_lib.REC_Z_F32 if f32_colour else _lib.REC_Z_U8)
    words = _lib.RECORD_BYTES[fmt] // 4
    with torch.cuda.device(dev):
        st = _stream(dev)
        masks = torch.empty((n_tiles, V), dtype=torch.int32, device=dev)
        view_count = torch.empty(V, dtype=torch.int64, device=dev)
        view_kept = torch.empty(V, dtype=torch.uint8, device=dev)
        rec_off = torch.empty(n_tiles + 1, dtype=torch.int64, device=dev)
        blk_off = torch.empty(n_tiles + 1, dtype=torch.int64, device=dev)
        row_off = torch.empty(n_tiles + 1, dtype=torch.int64, device=dev)
        totals = torch.zeros(5, dtype=torch.int64, device=dev)   # plan: {N, blocks, rows}; match statistics: {culled, in bounds}
        tptr = trec.ctypes.data
        _lib.check(L.sucre_gather_match(tptr, table.data_ptr(), V, C.byref(band), masks.data_ptr(),
                                        totals.data_ptr() + 24, st), 'sucre_gather_match')
        _lib.check(L.sucre_gather_count(masks.data_ptr(), n_tiles, V, view_count.data_ptr(), st), 'sucre_gather_count')
        if reduce_counts is not None:
            reduce_counts(view_count)
        if permute is None:
            permute = os.environ.get('SUCRE_PERMUTE', '1') != '0'
        pix = None
        if permute:
            pix = torch.empty(n_tiles * TILE, dtype=torch.int32, device=dev)
            pmasks = torch.empty_like(masks)
            _lib.check(L.sucre_gather_permute(masks.data_ptr(), V, C.byref(band), P, pix.data_ptr(), pmasks.data_ptr(), st),
                       'sucre_gather_permute')
            masks = pmasks
        _lib.check(L.sucre_gather_plan(masks.data_ptr(), n_tiles, V, view_count.data_ptr(), P, float(min_cover),
                                       view_kept.data_ptr(), rec_off.data_ptr(), blk_off.data_ptr(), row_off.data_ptr(),
                                       totals.data_ptr(), st), 'sucre_gather_plan')
        # the one host sync of the gather: the totals size the store; the per-view results ride along, so that nothing
        # after the sample launch makes the host wait for it
        host_side = torch.cat([totals, view_count, view_kept.to(torch.int64)]).cpu().numpy()
        n_obs, n_blocks, n_rows, culled, n_inb = (int(x) for x in host_side[:5])
        vc, vk = host_side[5:5 + V].copy(), host_side[5 + V:].astype(bool)
        cells = torch.empty((max(n_rows, 1), TILE, words), dtype=torch.float32, device=dev)
        blk_mask = torch.empty(max(n_blocks, 1), dtype=torch.int32, device=dev)
        blk_view = torch.empty(max(n_blocks, 1), dtype=torch.int32, device=dev)
        cell_src = torch.empty(max(n_rows, 1) * TILE, dtype=torch.int32, device=dev) if keep_src else None
        if n_obs > 0:
            missing = [k for k in source_keys if k not in scene.rgb]
            if missing:
                raise _lib.SucreError(f'gather: views without colour on the device: {missing[:3]}...')
            _lib.check(L.sucre_gather_sample(tptr, table.data_ptr(), V, C.byref(band), 0 if pix is None else pix.data_ptr(),
                                             masks.data_ptr(),
                                             view_kept.data_ptr(), row_off.data_ptr(), blk_off.data_ptr(), fmt,
                                             cells.data_ptr(), blk_mask.data_ptr(), blk_view.data_ptr(),
                                             0 if cell_src is None else cell_src.data_ptr(), st), 'sucre_gather_sample')
    stats = {'tile_views_culled': culled, 'tile_views': n_tiles * V, 'n_inbounds': n_inb}
    return ObservationStore(width=W, height=H, source_keys=source_keys, view_count=vc, view_kept=vk, n_obs=n_obs,
                            n_blocks=n_blocks, n_rows=n_rows, cells=cells, row_off=row_off, blk_off=blk_off, rec_off=rec_off,
                            blk_mask=blk_mask[:n_blocks], blk_view=blk_view[:n_blocks],
                            cell_src=None if cell_src is None else cell_src[:n_rows * TILE], band=band,
                            record_format=fmt, stats=stats, pix=pix)


# ------------------------------------------------------------------------------------------------------------
@dataclass
class FitState:
    """B, beta, gamma (9 floats, sucre.py:41-43), Adam's fp32 moments and the per-pixel J buffer, on the device.

    closed-form mode: J is the fit's work buffer (J of the last evaluated iteration on observed pixels, 0 elsewhere);
    J-parameter mode:  J is the Adam parameter (sucre.py:47-50) and J_moments its per-pixel {exp_avg, exp_avg_sq}."""
    params: torch.Tensor                  # (9,) f32: B[3], beta[3], gamma[3]
    moments: torch.Tensor                 # (18,) f32: exp_avg[9], exp_avg_sq[9]
    J: torch.Tensor | None = None         # (H,W,3) f32
    J_moments: torch.Tensor | None = None  # (H,W,6) f32
    step: int = 0

    @staticmethod
    def initial(device, params=None, J0: torch.Tensor | None = None) -> 'FitState':
        """J0 = None: closed-form mode; J0 = initial image (NaN where target depth <= 0): J-parameter mode."""
        p = torch.full((9,), 0.1, dtype=torch.float32) if params is None else \
            torch.as_tensor(params, dtype=torch.float32).reshape(9).clone()
        st = FitState(params=p.to(device), moments=torch.zeros(18, dtype=torch.float32, device=device))
        if J0 is not None:
            st.J = J0.to(device=device, dtype=torch.float32).contiguous().clone()
            st.J_moments = torch.zeros(tuple(st.J.shape[:-1]) + (6,), dtype=torch.float32, device=device)
        return st

    @property
    def mode(self) -> int:
        return _lib.FIT_PARAM_J if self.J_moments is not None else _lib.FIT_CLOSED_FORM

    def reset_optimizer(self):
        """A fresh Adam (what a new torch.optim.Adam starts from): step 0, zero moments; parameters and J are kept."""
        self.step = 0
        self.moments.zero_()
        if self.J_moments is not None:
            self.J_moments.zero_()

    def ensure_J(self, store: 'ObservationStore'):
        """closed-form mode: J is the loop's work buffer, in slot order; J-parameter mode: the caller's J (J_shape)."""
        if self.J is None:
            self.J = torch.zeros(store.slot_J_shape, dtype=torch.float32, device=store.cells.device)
        want = store.slot_J_shape if self.J_moments is None else store.J_shape
        assert tuple(self.J.shape) == want and self.J.is_contiguous(), (tuple(self.J.shape), want)

    def slot_arrays(self, store: 'ObservationStore'):
        """(J, J_moments or None, write_back) as the Adam-loop entry points address them.  Only the J-parameter mode on a
        permuted whole-image store needs copies: its J is the user's image-ordered parameter."""
        if self.J_moments is None or not store.needs_slot_copy:
            return self.J, self.J_moments, lambda: None
        J, Jm = store.to_slots(self.J), store.to_slots(self.J_moments)

        def write_back():
            store.from_slots(J, self.J)
            store.from_slots(Jm, self.J_moments)
        return J, Jm, write_back


def _workspace(store: ObservationStore) -> torch.Tensor:
    """Per-store scratch buffer, partitioned for the fit on first use."""
    if store.workspace is None:
        dev = store.cells.device
        store.workspace = torch.empty(_lib.lib().sucre_fit_workspace_bytes(), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            cs = store.c_struct()
            _lib.check(_lib.lib().sucre_fit_prepare(C.byref(cs), store.workspace.data_ptr(), _stream(dev)),
                       'sucre_fit_prepare')
    return store.workspace


def fit(store: ObservationStore, state: FitState, num_iter: int, lr: float = 0.05, peers=None,
        n_obs_global: int | None = None) -> torch.Tensor:
    """num_iter iterations of adam() (sucre.py:138-148) entirely on the device, in the mode `state` was created
    for.  Returns the (num_iter, 10) history tensor {params after each step, cost before it} (device).
    peers (a dist.PeerExchange): `store` is this rank's band of a target sharded over several GPUs; the all-reduce
    of the sums is then fused into the kernel over NVLink peer memory and n_obs_global is the all-band count."""
    if store.n_obs == 0 and peers is None:
        raise _lib.SucreError('fit: the observation store is empty')
    dev = store.cells.device
    state.ensure_J(store)
    history = torch.empty((num_iter, 10), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        J, Jm, write_back = state.slot_arrays(store)
        jm = 0 if Jm is None else Jm.data_ptr()
        cs = store.c_struct()
        if peers is None:
            _lib.check(_lib.lib().sucre_fit(
                state.mode, C.byref(cs), store.n_obs, state.params.data_ptr(), state.moments.data_ptr(),
                J.data_ptr(), jm, state.step + 1, num_iter, float(lr), history.data_ptr(),
                _workspace(store).data_ptr(), _stream(dev)), 'sucre_fit')
        else:
            ptrs = (C.c_uint64 * peers.world)(*peers.buffer_ptrs)
            _lib.check(_lib.lib().sucre_fit_sharded(
                state.mode, C.byref(cs), int(n_obs_global), state.params.data_ptr(), state.moments.data_ptr(),
                J.data_ptr(), jm, state.step + 1, num_iter, float(lr), history.data_ptr(),
                _workspace(store).data_ptr(), ptrs, peers.rank, peers.world, peers.take_epochs(num_iter), _stream(dev)),
                'sucre_fit_sharded')
        write_back()
    state.step += num_iter
    return history


def fit_status(store: ObservationStore) -> torch.Tensor:
    """Device uint32 with the status bits the fit kernels left in the store's workspace since the last call (0 = fine;
    bit 0: an in-kernel peer exchange timed out, i.e. a rank of a sharded fit went silent).  Enqueued on the current
    stream; reading it on the host synchronises."""
    dev = store.cells.device
    out = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().sucre_fit_status(_workspace(store).data_ptr(), out.data_ptr(), _stream(dev)), 'sucre_fit_status')
    return out


def scatter_J(store: ObservationStore, J_band: torch.Tensor, dst_ptrs: list[int]):
    """Writes a band's J (store.J_shape: slot order) to its place in whole-image J buffers given by device address — this
    GPU's or peers' NVLink-mapped ones (sucre_band_scatter_J)."""
    dev = store.cells.device
    ptrs = (C.c_uint64 * len(dst_ptrs))(*dst_ptrs)
    assert J_band.shape[0] == (store.n_slots if store.pix is not None else store.local_pixels)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().sucre_band_scatter_J(J_band.data_ptr(), C.byref(store.band),
                                                   0 if store.pix is None else store.pix.data_ptr(),
                                                   store.width * store.height, ptrs, len(dst_ptrs), _stream(dev)),
                   'sucre_band_scatter_J')


def fit_sums(store: ObservationStore, state: FitState, sums: torch.Tensor, n_obs_global: int | None = None,
             lr: float = 0.05):
    """One objective evaluation at state.params -> sums (10 doubles, device); in J-parameter mode J takes its
    Adam step state.step+1.  Building block of the multi-GPU loop (all-reduce sums, then adam_step)."""
    dev = store.cells.device
    state.ensure_J(store)
    with torch.cuda.device(dev):
        cs = store.c_struct()
        J, Jm, write_back = state.slot_arrays(store)
        _lib.check(_lib.lib().sucre_fit_sums(
            state.mode, C.byref(cs), state.params.data_ptr(), J.data_ptr(), 0 if Jm is None else Jm.data_ptr(),
            store.n_obs if n_obs_global is None else n_obs_global, state.step + 1, float(lr), sums.data_ptr(),
            _workspace(store).data_ptr(), _stream(dev)), 'sucre_fit_sums')
        write_back()


def adam_step(state: FitState, sums: torch.Tensor, n_obs: int, lr: float, history_row: torch.Tensor | None = None):
    dev = state.params.device
    state.step += 1
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().sucre_adam_step(state.params.data_ptr(), state.moments.data_ptr(), sums.data_ptr(), n_obs,
                                              state.step, float(lr), 0 if history_row is None else history_row.data_ptr(),
                                              _stream(dev)), 'sucre_adam_step')


def closed_form_J(store: ObservationStore, params: torch.Tensor, J_ref: torch.Tensor | None = None) -> torch.Tensor:
    """update_J (sucre.py:66-77) with the given parameters: store.J_shape f32 on the device, NaN where unobserved.
    J_ref: the work buffer of a closed-form fit (slot order), the reference point of the statistics."""
    dev = store.cells.device
    J = torch.empty(store.J_shape, dtype=torch.float32, device=dev)
    if store.n_obs == 0:
        return J.fill_(float('nan'))
    assert J_ref is None or tuple(J_ref.shape) == store.slot_J_shape, 'J_ref is the closed-form work buffer (slot order)'
    with torch.cuda.device(dev):
        cs = store.c_struct()
        _lib.check(_lib.lib().sucre_fit_write_J(C.byref(cs), params.data_ptr(), 0 if J_ref is None else J_ref.data_ptr(),
                                                J.data_ptr(), _workspace(store).data_ptr(), _stream(dev)),
                   'sucre_fit_write_J')
    return J


# ---- light model (sucre.py:44-46, 54-61) ----------------------------------------------------------------------
def light_J(store: ObservationStore, params24: torch.Tensor) -> torch.Tensor:
    """Closed-form J with the light terms for the 24 derived parameters (B, beta, gamma, R, t, Sigma^-1)."""
    dev = store.cells.device
    J = torch.empty(store.J_shape, dtype=torch.float32, device=dev)
    if store.n_obs == 0:
        return J.fill_(float('nan'))
    with torch.cuda.device(dev):
        cs = store.c_struct()
        _lib.check(_lib.lib().sucre_light_J(C.byref(cs), params24.data_ptr(), J.data_ptr(), _stream(dev)), 'sucre_light_J')
    return J


def light_sums(store: ObservationStore, params24: torch.Tensor, J: torch.Tensor, sums: torch.Tensor,
               J_moments: torch.Tensor | None = None, n_obs: int = 0, step: int = 0, lr: float = 0.05):
    """One residual pass -> sums (25 doubles, device).  With J_moments, J takes its Adam step `step` in the same pass."""
    dev = store.cells.device
    if store.workspace is None:
        store.workspace = torch.empty(_lib.lib().sucre_fit_workspace_bytes(), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        cs = store.c_struct()
        _lib.check(_lib.lib().sucre_light_sums(
            _lib.FIT_PARAM_J if J_moments is not None else _lib.FIT_CLOSED_FORM, C.byref(cs), params24.data_ptr(),
            J.data_ptr(), 0 if J_moments is None else J_moments.data_ptr(), n_obs, step, float(lr), sums.data_ptr(),
            store.workspace.data_ptr(), _stream(dev)), 'sucre_light_sums')
