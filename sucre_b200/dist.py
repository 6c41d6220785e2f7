"""Multi-GPU sharding of the hot path: one process per GPU, `torch.distributed` for the plumbing.

Two ways the path shards, both with the scene replicated on every GPU (a whole survey is <= 8.3 GB):

  * by target image   `shard_targets`: independent restorations, no data-path collective (what the CLI's
                      --image-list / --image-ids loops use under torchrun, configs 3 and 5);
  * by pixel band     `restore_band_sharded`: ONE target, every rank gathers and fits a contiguous band of its
                      tiles against all views (what `bench.py --gpus N` times, configs 2 and 4).  Exchange steps:
                      all-reduce(int64[V]) of the per-view match counts (min_cover is a whole-image criterion,
                      sfm.py:136; its kept-sum is the global N that normalises every gradient, sucre.py:135,145),
                      the reduction of the 10 residual sums once per Adam iteration (sucre.py:144-148) — fused into
                      the fit kernel over NVLink peer memory when a PeerExchange is given, an NCCL all-reduce between
                      two kernels otherwise — and the assembly of the J bands at the end (direct peer writes into
                      symmetric memory, or an all-gather).

The choreography is written against a small `ops` interface so that the same code runs with the CUDA kernels
(`CudaBandOps`, NCCL) and, in the CPU test-suite, with a numpy stand-in over gloo (tests/test_dist_gloo.py).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, engine
from ._lib import TILE


def shard_targets(targets: list, rank: int, world: int, contiguous: bool = False) -> list:
    """Assignment of target images to ranks: round-robin, or (contiguous) equal consecutive runs of the list — along a
    survey track neighbouring targets overlap the same source views, so a rank then decodes and keeps resident only
    its stretch of the survey."""
    if contiguous:
        n = len(targets)
        return list(targets[n * rank // world:n * (rank + 1) // world])
    return list(targets[rank::world])


def tile_band(n_tiles_total: int, rank: int, world: int) -> tuple[int, int]:
    """(first_tile, n_tiles) of `rank`: contiguous, balanced to within one tile, covering every tile once."""
    lo = n_tiles_total * rank // world
    hi = n_tiles_total * (rank + 1) // world
    return lo, hi - lo


def make_band(n_tiles_total: int, rank: int, world: int, layout: str = 'cyclic') -> _lib.Band:
    """The tiles of a target that `rank` of `world` restores.  'cyclic' (default): chunks of up to 64 tiles dealt
    round-robin, so that every rank sees every region of the image and the observation counts balance (contiguous
    bands differ by ~10 % on 8 ranks, tools/band_balance.py); 'contiguous': equal consecutive runs of tiles (what a rank
    that uploads from the host wants: its band then sees a small part of every source view)."""
    if world == 1:
        return _lib.Band.whole(n_tiles_total)
    if layout == 'contiguous':
        return _lib.Band.contiguous(n_tiles_total, rank, world)
    if layout != 'cyclic':
        raise ValueError(f'layout must be cyclic or contiguous, got {layout!r}')
    chunk = min(64, max(1, n_tiles_total // (world * 4)))
    return _lib.Band.cyclic(n_tiles_total, rank, world, chunk)


class PeerExchange:
    """Symmetric-memory buffers of a group of ranks (torch symmetric memory => every rank holds a device pointer to
    every peer's buffer, NVLink / NVSwitch peer access):
      * the exchange buffer of sucre_fit_sharded's in-kernel all-reduce (SUCRE_PEER_BUFFER_BYTES per rank), with the
        epoch tags, which must advance identically on all ranks;
      * on demand, a J buffer into which every rank writes its band directly (assemble_J), replacing the all-gather."""

    def __init__(self, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self._symm = symm_mem
        self.group = dist.group.WORLD if group is None else group
        self.device = torch.device(device)
        self.buffer = symm_mem.empty(_lib.PEER_BUFFER_BYTES // 4, dtype=torch.int32, device=device)
        self.buffer.zero_()
        self.handle = symm_mem.rendezvous(self.buffer, self.group)
        self.rank, self.world = self.handle.rank, self.handle.world_size
        if self.world > _lib.MAX_PEERS:
            raise _lib.SucreError(f'at most {_lib.MAX_PEERS} ranks per sharded target')
        self.buffer_ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self._epoch = 1
        self._J = None          # (capacity, 3) f32 symmetric
        self._J_handle = None
        torch.cuda.synchronize(device)
        dist.barrier(self.group)  # every buffer is zeroed before anybody writes into a peer

    def take_epochs(self, n: int) -> int:
        first = self._epoch
        self._epoch += n
        if self._epoch >= 2 ** 32:
            raise _lib.SucreError('epoch counter exhausted; create a new PeerExchange')
        return first

    def _ensure_J(self, pixels: int):
        if self._J is None or self._J.shape[0] < pixels:   # collective: every rank asks for the same size at the same call
            self._J = self._symm.empty((pixels, 3), dtype=torch.float32, device=self.device)
            self._J_handle = self._symm.rendezvous(self._J, self.group)

    def assemble_J(self, store: engine.ObservationStore, local: torch.Tensor, root_only: bool = False) -> torch.Tensor:
        """Every rank writes its band `local` (local_pixels, 3) to its place in the symmetric whole-image J buffer of
        every rank (or of rank 0 only) — one kernel storing through NVLink-mapped pointers (sucre_band_scatter_J) —
        then all ranks meet at a device-side barrier.  Returns this rank's (pixels, 3) buffer view: complete on every
        rank (on rank 0 only with root_only).  The buffer is overwritten by the next call: consume (or copy) the
        result on the same stream before restoring the next target."""
        pixels = store.width * store.height
        self._ensure_J(pixels)
        ptrs = [int(p) for p in self._J_handle.buffer_ptrs]
        engine.scatter_J(store, local.contiguous(), ptrs[:1] if root_only else ptrs)
        self._J_handle.barrier()   # stream-ordered after the stores: every band has landed everywhere
        return self._J[:pixels]


@dataclass
class BandResult:
    J: torch.Tensor          # (H,W,3) assembled on every rank
    params: torch.Tensor     # (9,)
    history: torch.Tensor    # (num_iter, 10)
    n_obs: int               # global
    view_kept: object
    status: torch.Tensor | None = None   # device uint32 (0 = fine; bit 0: a peer exchange timed out), fused path only
    n_local: int = 0


class CudaBandOps:
    """The CUDA kernels behind the band choreography (device tensors, NCCL-capable)."""

    def __init__(self, scene: engine.DeviceScene, target_key, source_keys, use_closed_form: bool = True):
        self.scene, self.target_key, self.source_keys = scene, target_key, tuple(source_keys)
        self.use_closed_form = use_closed_form
        self.device = scene.device
        rec = scene.record(target_key)
        self.width, self.height = int(rec['width']), int(rec['height'])
        self.store: engine.ObservationStore | None = None
        self.state: engine.FitState | None = None

    def gather(self, band: _lib.Band, min_cover, reduce_counts):
        self.store = engine.gather(self.scene, self.target_key, self.source_keys, min_cover=min_cover,
                                   band=band, reduce_counts=reduce_counts)
        return self.store.n_obs, self.store.view_kept, self.store.view_count

    def init_state(self, params=None):
        J0 = None
        if not self.use_closed_form:  # J parameter: this band of the target image, NaN where depth <= 0 (sucre.py:47-49)
            px = self.store.global_pixels()
            J0 = self.scene.rgb_float(self.target_key).reshape(-1, 3)[px]
            J0[self.scene.depth[self.target_key].view(torch.int16).reshape(-1)[px] == 0] = float('nan')
        self.state = engine.FitState.initial(self.device, params=params, J0=J0)
        self.state.ensure_J(self.store)

    def new_sums(self):
        return torch.zeros(10, dtype=torch.float64, device=self.device)

    def fit_sums(self, sums, n_obs_global, lr):
        engine.fit_sums(self.store, self.state, sums, n_obs_global=n_obs_global, lr=lr)

    def adam_step(self, sums, n_obs_global, lr, history_row):
        engine.adam_step(self.state, sums, n_obs_global, lr, history_row)

    def fused_fit(self, peers: 'PeerExchange', n_obs_global: int, num_iter: int, lr: float):
        """The whole Adam loop as one kernel per iteration with the all-reduce fused in (sucre_fit_sharded).  A band
        without observations still launches: it contributes zero sums and its epoch flag."""
        return engine.fit(self.store, self.state, num_iter, lr, peers=peers, n_obs_global=n_obs_global)

    def status(self):
        return engine.fit_status(self.store)

    def new_history(self, num_iter):
        return torch.empty((num_iter, 10), dtype=torch.float32, device=self.device)

    def band_pixels(self):
        """Flat image index of the pixel behind every row of band_J() (int64, -1: none)."""
        return self.store.global_pixels()

    def band_J(self):
        """J of this band (store.J_shape rows): closed form with the final parameters, or the optimised parameter."""
        if not self.use_closed_form:
            return self.state.J.reshape(-1, 3)
        return engine.closed_form_J(self.store, self.state.params, self.state.J).reshape(-1, 3)

    def params(self):
        return self.state.params


def restore_band_sharded(ops, *, min_cover: float = 1e-6, num_iter: int = 200, lr: float = 0.05, params=None,
                         group=None, peers: PeerExchange | None = None, root_only: bool = False,
                         layout: str = 'cyclic') -> BandResult:
    """One target restored by all ranks of `group`, each owning a band of its tiles (make_band).  Every rank returns the
    same parameters and the full J (with root_only and `peers`: J is complete on rank 0 only).  `ops` is a CudaBandOps
    (or a stand-in with the same methods).
    With `peers` (and a CudaBandOps) the per-iteration all-reduce runs inside the fit kernel over NVLink peer memory
    and the J bands are written straight into every rank's symmetric J buffer; without, both are NCCL / gloo
    collectives between kernels."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    P = ops.width * ops.height
    n_tiles_total = (P + TILE - 1) // TILE
    band = make_band(n_tiles_total, rank, world, layout)
    if band.n_tiles == 0:
        raise _lib.SucreError(f'restore: {world} ranks are too many for a target of {n_tiles_total} tiles')

    def all_reduce(t):
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)

    # 1. gather this band; min_cover decided on whole-image counts, whose kept-sum is the global n_obs (sucre.py:135)
    n_local, view_kept, view_count = ops.gather(band, min_cover, all_reduce)
    n_obs = int(np.asarray(view_count)[np.asarray(view_kept, dtype=bool)].sum())
    if n_obs == 0:
        raise _lib.SucreError('restore: no observation survives the two-way check and min_cover')

    # 2. Adam loop: local sums -> all-reduce(10 doubles) -> identical step on every rank
    ops.init_state(params)
    status = None
    fused = peers is not None and world > 1 and hasattr(ops, 'fused_fit')
    if fused:
        history = ops.fused_fit(peers, n_obs, num_iter, lr)
        status = ops.status()
    else:
        sums = ops.new_sums()
        history = ops.new_history(num_iter)
        for it in range(num_iter):
            ops.fit_sums(sums, n_obs, lr)
            all_reduce(sums)
            ops.adam_step(sums, n_obs, lr, history[it])

    # 3. assemble J
    local = ops.band_J()
    if fused:
        J = peers.assemble_J(ops.store, local, root_only=root_only)
    else:  # local sizes differ by at most one chunk: pad to the longest, gather J and every rank's pixel list, scatter
        rows = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
        if world > 1:
            dist.all_reduce(rows, op=dist.ReduceOp.MAX, group=group)
        longest = int(rows.item())
        if hasattr(ops, 'band_pixels'):
            px_local = ops.band_pixels()
        else:
            px_local = torch.from_numpy(band.pixels(P)).to(local.device)
        padded = torch.full((longest, 3), float('nan'), dtype=local.dtype, device=local.device)
        padded[:local.shape[0]] = local
        px_padded = torch.full((longest,), -1, dtype=torch.int64, device=local.device)
        px_padded[:px_local.shape[0]] = px_local
        if world > 1:
            parts = [torch.empty_like(padded) for _ in range(world)]
            px_parts = [torch.empty_like(px_padded) for _ in range(world)]
            dist.all_gather(parts, padded, group=group)
            dist.all_gather(px_parts, px_padded, group=group)
        else:
            parts, px_parts = [padded], [px_padded]
        J = torch.empty((P, 3), dtype=local.dtype, device=local.device)
        for px, part in zip(px_parts, parts):
            ok = px >= 0
            J[px[ok]] = part[ok]
    return BandResult(J=J.reshape(ops.height, ops.width, 3), params=ops.params(), history=history, n_obs=n_obs,
                      view_kept=view_kept, status=status, n_local=n_local)
