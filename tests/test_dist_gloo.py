"""N > 1 host logic on CPU: world_size-2 gloo run of the pixel-band choreography (sucre_b200/dist.py) with a
numpy/oracle stand-in for the CUDA kernels, checked against the single-process oracle."""
import os
import socket
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle
from sucre_b200 import dist as sdist
from sucre_b200._lib import TILE

from conftest import Golden


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


class OracleBandOps:
    """Same methods as dist.CudaBandOps, computed with the CPU oracle + numpy (test infrastructure)."""
    device = 'cpu'

    def __init__(self, g: Golden):
        self.g = g
        names = g['names'].tolist()
        self.geoms = {}
        for i in range(g.n_views):
            a = g.geom_arrays(i)
            self.geoms[names[i]] = oracle.view_geom(a['K'], a['R'], a['t'], a['wh'][0], a['wh'][1], Kinv=a['Kinv'],
                                                    Ri=a['Ri'], ti=a['ti'])
        self.names = names
        self.order = sorted(names[i] for i in g.pairing_list())
        self.target = str(g['target'])
        self.width, self.height = (int(x) for x in g.geom_arrays(g.view_index(self.target))['wh'])

    def _inputs(self, name):
        return self.g.inputs(self.g.view_index(name))

    def gather(self, band, min_cover, reduce_counts):
        P = self.width * self.height
        self.px = band.pixels(P)                     # global pixel of every local pixel
        self.local_of = np.full(P, -1, np.int64)
        self.local_of[self.px] = np.arange(len(self.px))
        mine = np.zeros(P, bool)
        mine[self.px] = True
        depthT = self._inputs(self.target)[0]
        per_view, counts = [], []
        for name in self.order:
            idx, _, _ = oracle.match_pair(depthT, self.geoms[self.target], self._inputs(name)[0], self.geoms[name])
            flat = idx.reshape(-1).copy()
            flat[~mine] = -1
            per_view.append(flat.reshape(idx.shape))
            counts.append(int((flat >= 0).sum()))
        view_count = torch.tensor(counts, dtype=torch.int64)
        reduce_counts(view_count)
        kept = (view_count.numpy() / (self.width * self.height)) > min_cover
        self.obs = []
        for name, idx, keep in zip(self.order, per_view, kept):
            if keep:
                depth, rgb = self._inputs(name)
                self.obs.append(oracle.sample_pair(idx, depth, rgb, self.geoms[name]))
        return sum(len(o['u1']) for o in self.obs), kept, view_count.numpy()

    def init_state(self, params=None):
        self.p = np.full(9, 0.1, np.float32) if params is None else np.asarray(params, np.float32)
        self.m = np.zeros(9, np.float32)
        self.v = np.zeros(9, np.float32)
        self.t = 0

    def new_sums(self):
        return torch.zeros(10, dtype=torch.float64)

    def new_history(self, n):
        return torch.zeros((n, 10), dtype=torch.float32)

    def _J(self):
        B, beta, gamma = self.p[0:3].astype(np.float64), self.p[3:6].astype(np.float64), self.p[6:9].astype(np.float64)
        n = len(self.px)
        num, den = np.zeros((n, 3)), np.zeros((n, 3))
        for o in self.obs:
            pix = self.local_of[o['v1'].astype(np.int64) * self.width + o['u1']]
            z = o['z'].astype(np.float64)[:, None]
            a = np.exp(-beta * z)
            num[pix] += (o['I'].T - B * (1 - np.exp(-gamma * z))) * a
            den[pix] += a * a
        with np.errstate(invalid='ignore', divide='ignore'):
            return num / den

    def fit_sums(self, sums, n_obs_global, lr):
        B, beta, gamma = self.p[0:3].astype(np.float64), self.p[3:6].astype(np.float64), self.p[6:9].astype(np.float64)
        J = self._J()
        out = np.zeros(10)
        for o in self.obs:
            pix = self.local_of[o['v1'].astype(np.int64) * self.width + o['u1']]
            z = o['z'].astype(np.float64)[:, None]
            a, e = np.exp(-beta * z), np.exp(-gamma * z)
            r = o['I'].T - (J[pix] * a + B * (1 - e))
            out[0:3] += (r * (1 - e)).sum(0)
            out[3:6] += (r * J[pix] * z * a).sum(0)
            out[6:9] += (r * B * z * e).sum(0)
            out[9] += (r * r).sum()
        sums.copy_(torch.from_numpy(out))

    def adam_step(self, sums, n_obs, lr, history_row):
        self.t += 1
        s = sums.numpy()
        sc = 2.0 / (3.0 * n_obs)
        g = np.concatenate([-sc * s[0:3], sc * s[3:6], -sc * s[6:9]]).astype(np.float32)
        f = np.float32
        self.m = self.m + f(1 - 0.9) * (g - self.m)
        self.v = self.v * f(0.999) + f(1 - 0.999) * (g * g)
        step = lr / (1 - 0.9 ** self.t)
        denom = np.sqrt(self.v) / f(np.sqrt(1 - 0.999 ** self.t)) + f(1e-8)
        self.p = (self.p + f(-step) * (self.m / denom)).astype(np.float32)
        history_row[:9] = torch.from_numpy(self.p)
        history_row[9] = float(s[9])

    def band_J(self):
        return torch.from_numpy(self._J().astype(np.float32))

    def params(self):
        return torch.from_numpy(self.p.copy())


class SlotOrderedBandOps(OracleBandOps):
    """Like a CudaBandOps on a permuted store: the band's J comes in SLOT order — the band's pixels shuffled, with a few
    slots that hold no pixel — and band_pixels() tells restore_band_sharded which image pixel every row belongs to."""

    def _slots(self):
        n = len(self.px)
        rng = np.random.default_rng(n)
        order = np.concatenate([rng.permutation(n), np.full(5, -1)])
        return order[rng.permutation(len(order))]

    def band_pixels(self):
        o = self._slots()
        return torch.from_numpy(np.where(o >= 0, self.px[np.maximum(o, 0)], -1).astype(np.int64))

    def band_J(self):
        o = self._slots()
        J = self._J().astype(np.float32)[np.maximum(o, 0)]
        J[o < 0] = 123.0   # a slot without a pixel: whatever it holds must never reach the image
        return torch.from_numpy(J)


def _worker(rank, world, port, case, out_path, slot_order=False):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = Golden(case)
        ops = SlotOrderedBandOps(g) if slot_order else OracleBandOps(g)
        res = sdist.restore_band_sharded(ops, min_cover=float(g['min_cover']), num_iter=int(g['num_iter']))
        gathered = [None] * world
        dist.all_gather_object(gathered, res.params.numpy().tolist())
        assert all(p == gathered[0] for p in gathered)  # every rank holds identical parameters
        if rank == 0:
            np.savez(out_path, J=res.J.numpy(), params=res.params.numpy(), history=res.history.numpy(),
                     n_obs=res.n_obs, kept=np.asarray(res.view_kept))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('case,world,slot_order', [('tiny6_closed', 2, False), ('mixed8_image0004', 2, False),
                                                   ('tiny6_closed', 3, False), ('tiny6_closed', 2, True)])
def test_band_sharded_restore_over_gloo(case, world, slot_order):
    """Cyclic bands (the default layout of restore_band_sharded) over gloo; slot_order: the bands' J rows come in the
    slot order of a permuted store, assembled through the gathered pixel lists."""
    g = Golden(case)
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, 'res.npz')
        mp.spawn(_worker, args=(world, _free_port(), case, out, slot_order), nprocs=world, join=True)
        z = np.load(out)
    # the reference's own result for the same scene
    ref_p = np.concatenate([g['B'].ravel(), g['beta'].ravel(), g['gamma'].ravel()])
    assert np.max(np.abs(z['params'] - ref_p) / np.abs(ref_p)) < 1e-4
    assert np.max(np.abs(z['history'][:, :9] - g['history']) / np.maximum(np.abs(g['history']), 0.05)) < 1e-4
    assert np.max(np.abs(z['history'][:, 9] - g['cost']) / g['cost']) < 2e-4
    assert int(z['n_obs']) == sum(len(g.matches(n)['u1']) for n in g['kept'].tolist())
    names = g['names'].tolist()
    order = sorted(names[i] for i in g.pairing_list())
    assert [n for n, k in zip(order, z['kept']) if k] == g['kept'].tolist()   # global min_cover decision
    assert np.array_equal(np.isnan(z['J']), np.isnan(g['J'])) and np.nanmax(np.abs(z['J'] - g['J'])) < 1e-4


def test_partitions_cover_everything_once():
    for n_tiles in (1, 7, 96, 38988, 259200):
        for world in (1, 2, 3, 4, 8):
            bands = [sdist.tile_band(n_tiles, r, world) for r in range(world)]
            assert bands[0][0] == 0 and sum(n for _, n in bands) == n_tiles
            assert all(bands[r][0] + bands[r][1] == bands[r + 1][0] for r in range(world - 1))
            assert max(n for _, n in bands) - min(n for _, n in bands) <= 1
            for layout in ('cyclic', 'contiguous'):      # every tile exactly once, in increasing order within a rank
                parts = [sdist.make_band(n_tiles, r, world, layout) for r in range(world)]
                tiles = [b.tiles() for b in parts]
                assert all((np.diff(t) > 0).all() for t in tiles)
                assert np.array_equal(np.sort(np.concatenate(tiles)), np.arange(n_tiles))
                P = n_tiles * 32 - 5
                px = np.concatenate([b.pixels(P) for b in parts])
                assert np.array_equal(np.sort(px), np.arange(P))
                if layout == 'cyclic' and n_tiles >= 96:
                    assert max(b.n_tiles for b in parts) - min(b.n_tiles for b in parts) <= 64
    targets = list(range(50))
    parts = [sdist.shard_targets(targets, r, 8) for r in range(8)]
    assert sorted(sum(parts, [])) == targets and max(map(len, parts)) - min(map(len, parts)) <= 1
