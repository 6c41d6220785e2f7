"""Multi-GPU sharding of the hot path: one process per GPU, `torch.distributed` for the plumbing.

Two ways the path shards, both with the scene replicated on every GPU (a whole survey is <= 8.3 GB):

  * by target image   `shard_targets`: independent restorations, no data-path collective (weak scaling; what
                      `bench.py --gpus N` and the CLI's --image-list / --image-ids loops use);
  * by pixel band     `restore_band_sharded`: ONE target, every rank gathers and fits a contiguous band of its
                      tiles against all views.  Collectives: all-reduce(int64[V]) of the per-view match counts
                      (min_cover is a whole-image criterion, sfm.py:136; its kept-sum is the global N that normalises
                      every gradient, sucre.py:135,145), all-reduce(f64[10]) of the residual sums once per Adam
                      iteration (sucre.py:144-148), all-gather of the J bands at the end.

The choreography is written against a small `ops` interface so that the same code runs with the CUDA kernels
(`CudaBandOps`, NCCL) and, in the CPU test-suite, with a numpy stand-in over gloo (tests/test_dist_gloo.py).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist

from . import _lib, engine
from ._lib import TILE


def shard_targets(targets: list, rank: int, world: int) -> list:
    """Round-robin assignment of target images to ranks (every rank keeps the whole scene)."""
    return list(targets[rank::world])


def tile_band(n_tiles_total: int, rank: int, world: int) -> tuple[int, int]:
    """(first_tile, n_tiles) of `rank`: contiguous, balanced to within one tile, covering every tile once."""
    lo = n_tiles_total * rank // world
    hi = n_tiles_total * (rank + 1) // world
    return lo, hi - lo


class PeerExchange:
    """Exchange buffers for the in-kernel all-reduce of sucre_fit_sharded: one SUCRE_PEER_BUFFER_BYTES buffer per rank
    in torch symmetric memory, so that every rank holds a device pointer to every peer's buffer (NVLink / NVSwitch
    peer access).  Also hands out the epoch tags, which must advance identically on all ranks."""

    def __init__(self, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        group = dist.group.WORLD if group is None else group
        self.buffer = symm_mem.empty(_lib.PEER_BUFFER_BYTES // 4, dtype=torch.int32, device=device)
        self.buffer.zero_()
        self.handle = symm_mem.rendezvous(self.buffer, group)
        self.rank, self.world = self.handle.rank, self.handle.world_size
        if self.world > _lib.MAX_PEERS:
            raise engine._lib.SucreError(f'at most {_lib.MAX_PEERS} ranks per sharded target')
        self.buffer_ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self._epoch = 1
        torch.cuda.synchronize(device)
        dist.barrier(group)  # every buffer is zeroed before anybody writes into a peer

    def take_epochs(self, n: int) -> int:
        first = self._epoch
        self._epoch += n
        if self._epoch >= 2 ** 32:
            raise engine._lib.SucreError('epoch counter exhausted; create a new PeerExchange')
        return first


@dataclass
class BandResult:
    J: torch.Tensor          # (H,W,3) assembled on every rank
    params: torch.Tensor     # (9,)
    history: torch.Tensor    # (num_iter, 10)
    n_obs: int               # global
    view_kept: object


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

    def gather(self, tile_range, min_cover, reduce_counts):
        self.store = engine.gather(self.scene, self.target_key, self.source_keys, min_cover=min_cover,
                                   tile_range=tile_range, reduce_counts=reduce_counts)
        return self.store.n_obs, self.store.view_kept

    def init_state(self, params=None):
        J0 = None
        if not self.use_closed_form:  # J parameter: this band of the target image, NaN where depth <= 0 (sucre.py:47-49)
            lo = self.store.first_tile * TILE
            hi = lo + self.store.local_pixels
            J0 = self.scene.rgb_float(self.target_key).reshape(-1, 3)[lo:hi].clone()
            J0[self.scene.depth[self.target_key].view(torch.int16).reshape(-1)[lo:hi] == 0] = float('nan')
        self.state = engine.FitState.initial(self.device, params=params, J0=J0)
        if self.store.n_obs > 0:
            self.state.ensure_J(self.store)
        elif self.state.J is None:
            self.state.J = torch.zeros(self.store.J_shape, dtype=torch.float32, device=self.device)

    def new_sums(self):
        return torch.zeros(10, dtype=torch.float64, device=self.device)

    def fit_sums(self, sums, n_obs_global, lr):
        if self.store.n_obs > 0:
            engine.fit_sums(self.store, self.state, sums, n_obs_global=n_obs_global, lr=lr)
        else:
            sums.zero_()

    def adam_step(self, sums, n_obs_global, lr, history_row):
        engine.adam_step(self.state, sums, n_obs_global, lr, history_row)

    def fused_fit(self, peers: 'PeerExchange', n_obs_global: int, num_iter: int, lr: float):
        """The whole Adam loop as one kernel per iteration with the all-reduce fused in (sucre_fit_sharded)."""
        return engine.fit(self.store, self.state, num_iter, lr, peers=peers, n_obs_global=n_obs_global)

    def new_history(self, num_iter):
        return torch.empty((num_iter, 10), dtype=torch.float32, device=self.device)

    def band_J(self):
        """(local_pixels, 3) J of this band: closed form with the final parameters, or the optimised parameter."""
        if not self.use_closed_form:
            return self.state.J.reshape(-1, 3)
        return engine.closed_form_J(self.store, self.state.params, self.state.J).reshape(-1, 3)

    def params(self):
        return self.state.params


def restore_band_sharded(ops, *, min_cover: float = 1e-6, num_iter: int = 200, lr: float = 0.05, params=None,
                         group=None, peers: PeerExchange | None = None) -> BandResult:
    """One target restored by all ranks of `group`, each owning a band of its pixels.  Every rank returns the same
    parameters and the full J.  `ops` is a CudaBandOps (or a stand-in with the same methods).
    With `peers` (and a CudaBandOps) the per-iteration all-reduce runs inside the fit kernel over NVLink peer memory;
    without, it is an NCCL / gloo all-reduce between a sums kernel and a step kernel."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    P = ops.width * ops.height
    n_tiles_total = (P + TILE - 1) // TILE
    band = tile_band(n_tiles_total, rank, world)

    def all_reduce(t):
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)

    # 1. gather this band; min_cover decided on whole-image counts
    n_local, view_kept = ops.gather(band, min_cover, all_reduce)
    n_obs = torch.tensor([n_local], dtype=torch.int64, device=ops.device)
    all_reduce(n_obs)
    n_obs = int(n_obs.item())
    if n_obs == 0:
        raise engine._lib.SucreError('restore: no observation survives the two-way check and min_cover')

    # 2. Adam loop: local sums -> all-reduce(10 doubles) -> identical step on every rank
    ops.init_state(params)
    if peers is not None and world > 1 and hasattr(ops, 'fused_fit'):
        history = ops.fused_fit(peers, n_obs, num_iter, lr)
    else:
        sums = ops.new_sums()
        history = ops.new_history(num_iter)
        for it in range(num_iter):
            ops.fit_sums(sums, n_obs, lr)
            all_reduce(sums)
            ops.adam_step(sums, n_obs, lr, history[it])

    # 3. assemble J: bands differ by at most one tile, pad to the longest
    local = ops.band_J()
    longest = (n_tiles_total + world - 1) // world * TILE
    padded = torch.full((longest, 3), float('nan'), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    if world > 1:
        parts = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(parts, padded, group=group)
    else:
        parts = [padded]
    J = torch.empty((P, 3), dtype=local.dtype, device=local.device)
    for r, part in enumerate(parts):
        lo, n = tile_band(n_tiles_total, r, world)
        lo_px, hi_px = lo * TILE, min(P, (lo + n) * TILE)
        J[lo_px:hi_px] = part[:hi_px - lo_px]
    return BandResult(J=J.reshape(ops.height, ops.width, 3), params=ops.params(), history=history, n_obs=n_obs,
                      view_kept=view_kept)
