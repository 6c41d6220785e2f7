"""Array-level entry points of the hot path (no files): what sucre.restore_image does between decode and save.

    restore_resident   scene already in HBM (engine.DeviceScene)      -> device tensors
    restore_from_host  scene in (pinned) host memory, copied H2D here -> host tensors

Both run: fused gather -> observation store -> Adam loop (closed form or J-parameter) -> final J, all as CUDA
kernels of libsucre_b200.so.  Replaces sfm.py:127-138 + loader.py:78-118 + sucre.py:124-157 of the reference.
"""
from __future__ import annotations

from dataclasses import dataclass

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

    def pin(self) -> 'HostScene':
        return HostScene(self.geoms, self.depth.pin_memory(), self.rgb.pin_memory())

    @property
    def nbytes(self) -> int:
        return self.depth.numel() * 2 + self.rgb.numel()


@dataclass
class RestoreResult:
    J: torch.Tensor          # (H,W,3) restored image, NaN where unobserved
    params: torch.Tensor     # (9,) B, beta, gamma
    history: torch.Tensor    # (num_iter, 10) params after each step + cost before it
    n_obs: int
    view_kept: object        # (V,) bool
    store: engine.ObservationStore | None = None
    state: engine.FitState | None = None


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


def restore_from_host(host: HostScene, target: int, sources=None, *, device='cuda', out_J: torch.Tensor | None = None,
                      **kw) -> RestoreResult:
    """End to end from host buffers: H2D of every listed view, restore, D2H of J, parameters and history.
    out_J: optional (H,W,3) float32 host tensor (ideally pinned) that receives J; otherwise a new pageable tensor."""
    sources = list(range(len(host.geoms))) if sources is None else list(sources)
    needed = sorted(set(sources) | {target})
    scene = engine.DeviceScene(device)
    if len(needed) == len(host.geoms):
        scene.add_views(needed, host.geoms, host.depth, host.rgb)
    else:
        for i in needed:
            scene.add_view(i, host.geoms[i], host.depth[i], host.rgb[i])
    res = restore_resident(scene, target, sources, **kw)
    if out_J is None:
        J = res.J.cpu()
    else:
        J = out_J.copy_(res.J, non_blocking=True)
    params, history = res.params.cpu(), res.history.cpu()  # synchronises the stream: J has landed too
    return RestoreResult(J=J, params=params, history=history, n_obs=res.n_obs, view_kept=res.view_kept)


def h2d_bytes(host: HostScene, target: int, sources=None) -> int:
    sources = list(range(len(host.geoms))) if sources is None else list(sources)
    n = len(set(sources) | {target})
    return n * (host.depth[0].numel() * 2 + host.rgb[0].numel())


def d2h_bytes(res: RestoreResult) -> int:
    return res.J.numel() * 4 + res.params.numel() * 4 + res.history.numel() * 4
