"""Array-level entry points of the hot path (no files): what sucre.restore_image does between decode and save.

    restore_resident          scene already in HBM (engine.DeviceScene)               -> device tensors
    restore_from_host         scene in (pinned) host memory, copied H2D here          -> host tensors
    restore_stream            a sequence of targets from host memory, double-buffered: the upload of target k+1 and
                              the read-back of target k-1 overlap the fit of target k -> host tensors
    restore_from_host_sharded ONE target over all ranks of a torchrun job, every rank uploading what its band needs

Both run: fused gather -> observation store -> Adam loop (closed form or J-parameter) -> final J, all as CUDA
kernels of libsucre_b200.so.  Replaces sfm.py:127-138 + loader.py:78-118 + sucre.py:124-157 of the reference.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import torch

from . import engine

# kernels launched per restored image, for bench.py's `gpu_launches` claim: gather_match, count_views, permute, tile_count,
# scan, gather_sample, partition, fit_kernel (ONE resident launch runs all the Adam iterations), fit_kernel<write J>
LAUNCHES_PER_IMAGE = 9
# a band of a sharded target adds: status kernel, scatter_J (the device-side barrier and the NCCL all-reduce of the
# view counts are library kernels and not counted)
LAUNCHES_PER_BAND = LAUNCHES_PER_IMAGE + 2


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
    (the target's depth range comes from the host copy, so no device round trip is needed to plan the rectangles).
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
        needed, rects = upload_plan(host, target, sources, upload)
        d, c = scene.allocate_views(needed, [host.geoms[i] for i in needed], host.depth.dtype)
        h2d = scene.upload_rects((d, c), host.depth, host.rgb, needed, rects)
    res = restore_resident(scene, target, sources, **kw)
    if out_J is None:
        J = res.J.cpu()
    else:
        J = out_J.copy_(res.J, non_blocking=True)
    params, history = res.params.cpu(), res.history.cpu()  # synchronises the stream: J has landed too
    return RestoreResult(J=J, params=params, history=history, n_obs=res.n_obs, view_kept=res.view_kept, h2d_bytes=h2d)


def restore_stream(host: HostScene, targets, sources=None, *, device='cuda', upload: str = 'footprint', out_J=None, **kw):
    """Generator: restores `targets` one after the other from host buffers and yields a RestoreResult (host tensors)
    per target, in order.  Two device scene buffers and a copy stream: while target k is being fitted, the planes of
    target k+1 are already crossing PCIe into the other buffer and J of target k-1 is on its way back, so a
    multi-target run (configs 3 and 5) is bound by the kernels, not by the copies.  Every target's inputs are copied
    again — nothing is reused between targets — and the results are those of restore_from_host bit for bit.
    out_J: optional list of two pinned (H,W,3) float32 tensors the results alternate between (a yielded J is valid
    until the next-but-one target is yielded); otherwise every result gets its own pageable tensor."""
    targets = list(targets)
    sources = list(range(len(host.geoms))) if sources is None else list(sources)
    dev = torch.device(device)
    if not targets:
        return
    needed = sorted(set(sources) | set(targets))
    geoms = [host.geoms[i] for i in needed]
    assert all(g.width == geoms[0].width and g.height == geoms[0].height for g in geoms), 'restore_stream needs equally sized views'
    slot_of = {v: i for i, v in enumerate(needed)}
    copy_stream = torch.cuda.Stream(dev)
    compute = torch.cuda.current_stream(dev)
    scenes, planes = [], []
    for _ in range(2):  # depth planes start as zeros (= invalid); a later target overwrites only its own rectangles, and the
        sc = engine.DeviceScene(dev)   # gather never reads outside them, so what earlier targets left elsewhere is never seen
        planes.append(sc.allocate_views(needed, geoms, host.depth.dtype))
        scenes.append(sc)
    uploaded = [torch.cuda.Event() for _ in targets]
    consumed = [torch.cuda.Event() for _ in targets]   # the scene buffer of target k may be overwritten
    h2d = [0] * len(targets)

    def start_upload(k):
        t = targets[k]
        if upload == 'full':
            rects = np.array([[0, 0, g.width, g.height] for g in geoms], dtype=np.int32)
        else:
            views, r = upload_plan(host, t, sources, upload)
            rects = np.zeros((len(needed), 4), dtype=np.int32)
            rects[[slot_of[v] for v in views]] = r
        with torch.cuda.stream(copy_stream):
            if k >= 2:
                copy_stream.wait_event(consumed[k - 2])
            h2d[k] = scenes[k % 2].upload_rects(planes[k % 2], host.depth, host.rgb, needed, rects)
            uploaded[k].record(copy_stream)

    pending = None   # (result with device tensors, host J, readback event) of the previous target
    start_upload(0)
    for k, t in enumerate(targets):
        if k + 1 < len(targets):
            start_upload(k + 1)
        compute.wait_event(uploaded[k])
        res = restore_resident(scenes[k % 2], t, sources, **kw)   # its one host sync (store size) also paces this loop
        consumed[k].record(compute)
        if pending is not None:
            yield _finish(pending)
        J_host = out_J[k % 2] if out_J is not None else torch.empty(tuple(res.J.shape), dtype=torch.float32).pin_memory()
        small_dev = torch.cat([res.params, res.history.reshape(-1)])
        small = torch.empty(small_dev.shape, dtype=torch.float32).pin_memory()
        ready, done = torch.cuda.Event(), torch.cuda.Event()
        ready.record(compute)
        with torch.cuda.stream(copy_stream):   # read-back on the copy stream: the next gather does not wait for it
            copy_stream.wait_event(ready)
            J_host.copy_(res.J, non_blocking=True)
            small.copy_(small_dev, non_blocking=True)
            done.record(copy_stream)
        res.J.record_stream(copy_stream)
        small_dev.record_stream(copy_stream)
        pending = (res, J_host, small, done, h2d[k])
    yield _finish(pending)


def _finish(pending) -> RestoreResult:
    res, J_host, small, done, h2d = pending
    done.synchronize()
    return RestoreResult(J=J_host, params=small[:9].clone(), history=small[9:].reshape(-1, 10).clone(), n_obs=res.n_obs,
                         view_kept=res.view_kept, h2d_bytes=h2d)


def _band_rects(host: HostScene, target: int, needed: list, rank: int, world: int, upload: str) -> np.ndarray:
    """Per view of `needed`, the rectangle rank `rank` of `world` must hold to restore its (contiguous) band of `target`:
    the band's rows of the target itself and the footprint of that band in every view."""
    from . import dist as sdist
    g = host.geoms[target]
    if upload == 'full':
        return np.array([[0, 0, host.geoms[i].width, host.geoms[i].height] for i in needed], dtype=np.int32)
    n_tiles = (g.width * g.height + engine.TILE - 1) // engine.TILE
    lo, n = sdist.tile_band(n_tiles, rank, world)   # contiguous bands: a rank then needs a small part of every source view
    v0, v1 = lo * engine.TILE // g.width, min(g.height, ((lo + n) * engine.TILE - 1) // g.width + 1)
    rects = engine.DeviceScene.footprints(g, host_depth_range(host.depth[target]), None, rows_only=upload == 'rows',
                                          stacks=host.projection_stacks(needed), band_rows=(v0, v1))
    t = needed.index(target)   # the target is read as the target (its band's rows) and as a source (its footprint)
    ft = rects[t]
    rects[t] = [0, min(v0, int(ft[1])) if ft[3] > ft[1] else v0, g.width, max(v1, int(ft[3]))]
    return rects


def restore_stream_sharded(host: HostScene, targets, sources=None, *, device='cuda', peers=None, upload: str = 'footprint',
                           out_J=None, min_cover: float = 1e-6, use_closed_form: bool = True, num_iter: int = 200,
                           lr: float = 0.05, params=None):
    """restore_stream for targets sharded over all ranks of the default process group (a COLLECTIVE generator: every rank
    iterates it in step).  Per rank two device scene buffers and a copy stream: while the ranks fit their bands of target
    k, every rank's rectangles for target k+1 are already crossing PCIe and rank 0 reads J of target k-1 back.  Yields a
    RestoreResult per target (J on rank 0 only), the same as restore_from_host_sharded's bit for bit.
    out_J: optional list of two pinned (H,W,3) float32 tensors rank 0's results alternate between."""
    import torch.distributed as tdist
    from . import dist as sdist
    targets = list(targets)
    if not targets:
        return
    sources = list(range(len(host.geoms))) if sources is None else list(sources)
    world = tdist.get_world_size() if tdist.is_initialized() else 1
    rank = tdist.get_rank() if tdist.is_initialized() else 0
    dev = torch.device(device)
    needed = sorted(set(sources) | set(targets))
    geoms = [host.geoms[i] for i in needed]
    assert all(g.width == geoms[0].width and g.height == geoms[0].height for g in geoms), 'restore_stream_sharded needs equally sized views'
    copy_stream = torch.cuda.Stream(dev)
    compute = torch.cuda.current_stream(dev)
    scenes, planes = [], []
    for _ in range(2):
        sc = engine.DeviceScene(dev)
        planes.append(sc.allocate_views(needed, geoms, host.depth.dtype))
        scenes.append(sc)
    uploaded = [torch.cuda.Event() for _ in targets]
    consumed = [torch.cuda.Event() for _ in targets]
    h2d = [0] * len(targets)

    def start_upload(k):
        rects = _band_rects(host, targets[k], needed, rank, world, upload)
        with torch.cuda.stream(copy_stream):
            if k >= 2:
                copy_stream.wait_event(consumed[k - 2])
            h2d[k] = scenes[k % 2].upload_rects(planes[k % 2], host.depth, host.rgb, needed, rects)
            uploaded[k].record(copy_stream)

    def finish(pending):
        res, J_host, small, done, nbytes = pending
        done.synchronize()
        if small[-1] != 0:
            raise engine._lib.SucreError('restore_stream_sharded: a peer went silent during the in-kernel all-reduce')
        return RestoreResult(J=J_host, params=small[:9].clone(), history=small[9:-1].reshape(-1, 10).clone(), n_obs=res.n_obs,
                             view_kept=res.view_kept, h2d_bytes=nbytes)

    pending = None
    start_upload(0)
    for k, t in enumerate(targets):
        if k + 1 < len(targets):
            start_upload(k + 1)
        compute.wait_event(uploaded[k])
        ops = sdist.CudaBandOps(scenes[k % 2], t, sources, use_closed_form=use_closed_form)
        res = sdist.restore_band_sharded(ops, min_cover=min_cover, num_iter=num_iter, lr=lr, params=params, peers=peers,
                                         root_only=peers is not None, layout='contiguous')
        consumed[k].record(compute)
        if pending is not None:
            yield finish(pending)
        status = torch.zeros(1, device=dev) if res.status is None else res.status.to(torch.float32).reshape(1)
        small_dev = torch.cat([res.params, res.history.reshape(-1), status])
        small = torch.empty(small_dev.shape, dtype=torch.float32).pin_memory()
        J_host = J_dev = None
        if rank == 0:
            # the assembled J lives in the symmetric buffer every rank writes the NEXT target into: a device copy (microseconds)
            # is what crosses PCIe while the next target is fitted
            J_dev = res.J.clone()
            J_host = out_J[k % 2] if out_J is not None else torch.empty(tuple(J_dev.shape), dtype=torch.float32).pin_memory()
        ready, done = torch.cuda.Event(), torch.cuda.Event()
        ready.record(compute)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ready)
            if J_dev is not None:
                J_host.copy_(J_dev, non_blocking=True)
                J_dev.record_stream(copy_stream)
            small.copy_(small_dev, non_blocking=True)
            done.record(copy_stream)
        small_dev.record_stream(copy_stream)
        pending = (res, J_host, small, done, h2d[k])
    yield finish(pending)


def restore_from_host_sharded(host: HostScene, target: int, sources=None, *, device='cuda', peers=None,
                              out_J: torch.Tensor | None = None, upload: str = 'footprint', min_cover: float = 1e-6,
                              use_closed_form: bool = True, num_iter: int = 200, lr: float = 0.05, params=None) -> RestoreResult:
    """ONE target restored by all ranks of the default process group (dist.restore_band_sharded), end to end from host
    buffers: every rank copies to its GPU only what ITS band of the target can see (the band's rows of the target and
    the footprint rectangle of that band in every source view), restores its band, and rank 0 reads J, the
    parameters and the history back.  peers: a dist.PeerExchange (in-kernel all-reduce + direct J assembly)."""
    import torch.distributed as tdist
    from . import dist as sdist
    sources = list(range(len(host.geoms))) if sources is None else list(sources)
    world = tdist.get_world_size() if tdist.is_initialized() else 1
    rank = tdist.get_rank() if tdist.is_initialized() else 0
    needed = sorted(set(sources) | {target})
    scene = engine.DeviceScene(device)
    rects = _band_rects(host, target, needed, rank, world, upload)
    d, c = scene.allocate_views(needed, [host.geoms[i] for i in needed], host.depth.dtype)
    h2d = scene.upload_rects((d, c), host.depth, host.rgb, needed, rects)
    ops = sdist.CudaBandOps(scene, target, sources, use_closed_form=use_closed_form)
    res = sdist.restore_band_sharded(ops, min_cover=min_cover, num_iter=num_iter, lr=lr, params=params, peers=peers,
                                     root_only=peers is not None, layout='contiguous')
    J = None
    if rank == 0:
        J = res.J.cpu() if out_J is None else out_J.copy_(res.J, non_blocking=True)
    p, h = res.params.cpu(), res.history.cpu()  # synchronises the stream
    if res.status is not None and int(res.status.item()) != 0:
        raise engine._lib.SucreError('restore_from_host_sharded: a peer went silent during the in-kernel all-reduce')
    return RestoreResult(J=J, params=p, history=h, n_obs=res.n_obs, view_kept=res.view_kept, h2d_bytes=h2d)


def h2d_bytes(host: HostScene, target: int, sources=None, upload: str = 'full') -> int:
    """Bytes restore_from_host copies to the device for this call."""
    sources = list(range(len(host.geoms))) if sources is None else list(sources)
    _, rects = upload_plan(host, target, sources, upload)
    return int(((rects[:, 2] - rects[:, 0]).astype(np.int64) * (rects[:, 3] - rects[:, 1])).sum()) * 5


def d2h_bytes(res: RestoreResult) -> int:
    return res.J.numel() * 4 + res.params.numel() * 4 + res.history.numel() * 4
