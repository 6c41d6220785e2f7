"""Array-level entry points of the hot path (no files): what sucre.restore_image does between decode and save.

    restore_resident   scene already in HBM (engine.DeviceScene)      -> device tensors
    restore_from_host  scene in (pinned) host memory, copied H2D here -> host tensors

Both run: fused gather -> observation store -> Adam loop (closed form or J-parameter) -> final J, all as CUDA
kernels of libsucre_b200.so.  Replaces sfm.py:127-138 + loader.py:78-118 + sucre.py:124-157 of the reference.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch

from . import engine

# kernels launched per restored image, for bench.py's `gpu_launches` claim:
# gather_match, count_views, kept, tile_count, scan, gather_sample, partition, num_iter x fit_kernel, write_J
LAUNCHES_FIXED = 8


@dataclass
class HostScene:
    """A decoded survey in host memory: stacked u16 depth (V,H,W), u8 colour (V,H,W,3), one ViewGeom per view."""
    geoms: list
    depth: torch.Tensor
    rgb: torch.Tensor
    _stacks: tuple | None = field(default=None, repr=False, compare=False)

    def pin(self) -> 'HostScene':
        return HostScene(self.geoms, self.depth.pin_memory(), self.rgb.pin_memory())

    @property
    def nbytes(self) -> int:
        return self.depth.numel() * 2 + self.rgb.numel()

    def projection_stacks(self, views) -> tuple:
        """engine.projection_stacks of `views` (float64 copies of the per-view constants, converted once per scene)."""
        if self._stacks is None:
            self._stacks = engine.projection_stacks(self.geoms)
        idx = np.asarray(views, dtype=np.int64)
        return tuple(a[idx] for a in self._stacks)


@dataclass
class RestoreResult:
    J: torch.Tensor          # (H,W,3) restored image, NaN where unobserved
    params: torch.Tensor     # (9,) B, beta, gamma
    history: torch.Tensor    # (num_iter, 10) params after each step + cost before it
    n_obs: int
    view_kept: object        # (V,) bool
    store: engine.ObservationStore | None = None
    state: engine.FitState | None = None
    h2d_bytes: int = 0       # bytes copied host -> device by restore_from_host


def restore_resident(scene: engine.DeviceScene, target_key, source_keys, *, min_cover: float = 1e-6,
                     use_closed_form: bool = True, num_iter: int = 200, lr: float = 0.05, params=None,
                     keep_src: bool = False) -> RestoreResult:
    store = engine.gather(scene, target_key, source_keys, min_cover=min_cover, keep_src=keep_src)
    dev = scene.device
    J0 = None
    if not use_closed_form:  # sucre.py:47-49: J starts as the target image, NaN where its depth <= 0
        J0 = scene.rgb_float(target_key)
        J0[scene.depth[target_key].view(torch.int16) == 0] = float('nan')
    state = engine.FitState.initial(dev, params=params, J0=J0)
    if store.n_obs == 0:
        raise engine._lib.SucreError('restore: no observation survives the two-way check and min_cover')
    history = engine.fit(store, state, num_iter, lr)
    J = engine.closed_form_J(store, state.params, state.J) if use_closed_form else state.J
    return RestoreResult(J=J, params=state.params, history=history, n_obs=store.n_obs, view_kept=store.view_kept,
                         store=store, state=state)


UPLOAD_MODES = ('footprint', 'rows', 'full')


def host_depth_range(depth_u16: torch.Tensor) -> tuple[float, float]:
    """(smallest non-zero, largest) depth in metres of a host u16 millimetre plane; (65.536, 0.0) if it has no valid pixel."""
    a = depth_u16.numpy().view(np.uint16)
    lo = int((a - np.uint16(1)).min()) + 1  # 0 wraps to 65535: the minimum skips invalid pixels
    return lo / 1000.0, int(a.max()) / 1000.0


def upload_plan(host: HostScene, target: int, sources, upload: str = 'footprint', depth_range=None):
    """Which part of which host view a restoration of `target` against `sources` needs on the device:
    (needed view indices, (n,4) int32 rectangles x0, y0, x1, y1).  'full': whole views; 'footprint': the target whole,
    of every other view the rectangle the target can see (engine.DeviceScene.footprints); 'rows': the same widened to
    whole rows.  depth_range: (smallest non-zero, largest) target depth in metres if the caller already has it."""
    if upload not in UPLOAD_MODES:
        raise ValueError(f'upload must be one of {UPLOAD_MODES}, got {upload!r}')
    needed = sorted(set(sources) | {target})
    rects = np.array([[0, 0, host.geoms[i].width, host.geoms[i].height] for i in needed], dtype=np.int32)
    if upload != 'full':
        rng = host_depth_range(host.depth[target]) if depth_range is None else depth_range
        fp = engine.DeviceScene.footprints(host.geoms[target], rng, None, rows_only=upload == 'rows',
                                           stacks=host.projection_stacks(needed))
        fp[needed.index(target)] = rects[needed.index(target)]
        rects = fp
    return needed, rects


def restore_from_host(host: HostScene, target: int, sources=None, *, device='cuda', out_J: torch.Tensor | None = None,
                      upload: str = 'footprint', **kw) -> RestoreResult:
    """End to end from host buffers: H2D of what the listed views contribute, restore, D2H of J, parameters and
    history.  upload: see upload_plan — every mode gives the same result bit for bit, the footprint modes copy less
    (the target goes first and whole, its depth range is reduced on the device, then the rectangles follow).
    out_J: optional (H,W,3) float32 host tensor (ideally pinned) that receives J; otherwise a new pageable tensor."""
    sources = list(range(len(host.geoms))) if sources is None else list(sources)
    scene = engine.DeviceScene(device)
    if upload == 'full':
        needed, _ = upload_plan(host, target, sources, 'full')
        if len(needed) == len(host.geoms):
            scene.add_views(needed, host.geoms, host.depth, host.rgb)
        else:
            for i in needed:
                scene.add_view(i, host.geoms[i], host.depth[i], host.rgb[i])
        h2d = len(needed) * (host.depth[0].numel() * 2 + host.rgb[0].numel())
    else:
        needed = sorted(set(sources) | {target})
        d, c = scene.allocate_views(needed, [host.geoms[i] for i in needed], host.depth.dtype)
        t = needed.index(target)
        g = host.geoms[target]
        h2d = scene.upload_rects((d[t:t + 1], c[t:t + 1]), host.depth, host.rgb, [target], [[0, 0, g.width, g.height]])
        _, rects = upload_plan(host, target, sources, upload, depth_range=scene.depth_range(target))
        rects[t] = 0  # already there
        h2d += scene.upload_rects((d, c), host.depth, host.rgb, needed, rects)
    res = restore_resident(scene, target, sources, **kw)
    if out_J is None:
        J = res.J.cpu()
    else:
        J = out_J.copy_(res.J, non_blocking=True)
    params, history = res.params.cpu(), res.history.cpu()  # synchronises the stream: J has landed too
    return RestoreResult(J=J, params=params, history=history, n_obs=res.n_obs, view_kept=res.view_kept, h2d_bytes=h2d)


def h2d_bytes(host: HostScene, target: int, sources=None, upload: str = 'full') -> int:
    """Bytes restore_from_host copies to the device for this call."""
    sources = list(range(len(host.geoms))) if sources is None else list(sources)
    _, rects = upload_plan(host, target, sources, upload)
    return int(((rects[:, 2] - rects[:, 0]).astype(np.int64) * (rects[:, 3] - rects[:, 1])).sum()) * 5


def d2h_bytes(res: RestoreResult) -> int:
    return res.J.numel() * 4 + res.params.numel() * 4 + res.history.numel() * 4
