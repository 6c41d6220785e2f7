"""Pixel-band sharding on the GPU: bands compose to the whole-image result (1 GPU), and the NCCL choreography
reproduces the single-GPU restoration (>= 2 GPUs, skipped otherwise)."""
import os
import socket
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sucre_b200 import api, engine
from sucre_b200 import dist as sdist
from sucre_b200.synth import SyntheticScene

import helpers

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('layout', ['contiguous', 'cyclic'])
def test_bands_compose_to_the_whole_image(layout):
    scene = SyntheticScene(7, 150, 101, seed=9)   # 15150 pixels = 473 tiles + 14 pixels
    ds, _ = helpers.build_device_scene(scene, range(7))
    keys = list(range(7))
    full = engine.gather(ds, 3, keys, min_cover=0.2, keep_src=True)
    n_tiles = full.n_tiles
    from sucre_b200._lib import Band
    counts = []
    for r in range(3):
        counts.append(engine.gather(ds, 3, keys, min_cover=0.0, band=sdist.make_band(n_tiles, r, 3, layout)).view_count)
    total = torch.from_numpy(np.sum(counts, axis=0)).cuda()
    assert np.array_equal(total.cpu().numpy(), full.view_count)
    bands = [engine.gather(ds, 3, keys, min_cover=0.2, keep_src=True, band=sdist.make_band(n_tiles, r, 3, layout),
                           reduce_counts=lambda vc: vc.copy_(total)) for r in range(3)]
    assert sum(b.n_obs for b in bands) == full.n_obs and all(np.array_equal(b.view_kept, full.view_kept) for b in bands)
    assert sum(int((b.global_pixels() >= 0).sum()) for b in bands) == 150 * 101
    whole = full.to_reference_layout()
    parts = [b.to_reference_layout() for b in bands]
    for key, ref in whole.items():
        order = np.argsort(np.concatenate([p[key]['v1'].astype(np.int64) * 150 + p[key]['u1'] for p in parts]), kind='stable')
        for f in ('u1', 'v1', 'u2', 'v2', 'z', 'I'):
            cat = np.concatenate([p[key][f] for p in parts], axis=-1)
            assert np.array_equal(cat[..., order], ref[f]), (key, f)   # the bands' records, in target order, are the image's
    # one objective evaluation: band sums add up to the whole-image sums; band J's tile the whole J
    state = engine.FitState.initial(ds.device)
    sums = torch.zeros(10, dtype=torch.float64, device=ds.device)
    engine.fit_sums(full, state, sums)
    acc = torch.zeros_like(sums)
    Js = []
    for b in bands:
        sb = engine.FitState.initial(ds.device)
        s = torch.zeros_like(sums)
        engine.fit_sums(b, sb, s)
        acc += s
        Js.append(engine.closed_form_J(b, sb.params))
    # band sums add up to the whole-image sums (not bit for bit: the fit cuts tiles between warps at different places
    # in a band and in the whole image, which reorders the fp32 additions inside a pixel's statistics)
    assert torch.allclose(acc, sums, rtol=2e-4, atol=0)
    J = engine.closed_form_J(full, state.params).reshape(-1, 3)
    Jb = torch.full_like(J, -5.0)
    for b, Jr in zip(bands, Js):
        px = b.global_pixels()
        Jb[px[px >= 0]] = Jr[px >= 0]
    assert torch.equal(torch.isnan(Jb), torch.isnan(J))
    assert float((Jb - J).nan_to_num(0.0).abs().max()) < 1e-6
    # the scatter kernel puts a band's J at the same places
    out = torch.full_like(J, -5.0)
    for b, Jr in zip(bands, Js):
        engine.scatter_J(b, Jr.contiguous(), [out.data_ptr()])
    assert torch.equal(out.nan_to_num(7.0), Jb.nan_to_num(7.0))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _blank_top(ds, world):
    """Invalidate the target's first band (and a bit more): rank 0's band then holds no observation at all."""
    ds.depth[4].view(torch.int16)[:240 // world + 2] = 0


def _nccl_worker(rank, world, port, closed_form, out_path, fused=False, empty_band=False):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        scene = SyntheticScene(8, 320, 240, seed=11)
        ds, _ = helpers.build_device_scene(scene, range(8), device=f'cuda:{rank}')
        if empty_band:
            _blank_top(ds, world)
        ops = sdist.CudaBandOps(ds, 4, list(range(8)), use_closed_form=closed_form)
        peers = sdist.PeerExchange(ds.device) if fused else None
        res = sdist.restore_band_sharded(ops, num_iter=25, peers=peers)
        assert res.n_local > 0 or (empty_band and rank == 0)
        J = res.J.clone()   # the fused path returns a view of the symmetric buffer, overwritten by the next target
        if fused:  # a second target on the same exchange buffers: epochs keep advancing; no exchange timed out
            assert int(res.status.item()) == 0
            ops2 = sdist.CudaBandOps(ds, 3, list(range(8)), use_closed_form=closed_form)
            res2 = sdist.restore_band_sharded(ops2, num_iter=5, peers=peers, root_only=True)
            assert torch.isfinite(res2.params).all() and int(res2.status.item()) == 0
            full = torch.tensor([int(torch.isfinite(J).any())], device=ds.device)
            dist.all_reduce(full)
            assert int(full.item()) == world   # every rank holds the assembled J of the first target
        if rank == 0:
            np.savez(out_path, J=J.cpu().numpy(), params=res.params.cpu().numpy(), history=res.history.cpu().numpy(),
                     n_obs=res.n_obs)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
@pytest.mark.parametrize('closed_form,fused,empty_band', [(True, False, False), (False, False, False), (True, True, False),
                                                          (False, True, False), (True, True, True), (True, False, True)])
def test_band_sharded_matches_single_gpu(closed_form, fused, empty_band):
    """fused=False: NCCL all-reduce between kernels; fused=True: all-reduce inside the fit kernel over NVLink peer
    memory (sucre_fit_sharded) and J assembled by direct peer writes.  empty_band: the first rank's band has no
    observation at all — it must still take part in every exchange (a missing flag would hang its peers)."""
    world = min(4, torch.cuda.device_count())
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, 'res.npz')
        mp.spawn(_nccl_worker, args=(world, _free_port(), closed_form, out, fused, empty_band), nprocs=world, join=True)
        z = np.load(out)
    scene = SyntheticScene(8, 320, 240, seed=11)
    ds, _ = helpers.build_device_scene(scene, range(8))
    if empty_band:
        _blank_top(ds, world)
    one = api.restore_resident(ds, 4, list(range(8)), use_closed_form=closed_form, num_iter=25)
    assert int(z['n_obs']) == one.n_obs
    p1 = one.params.cpu().numpy()
    assert np.max(np.abs(z['params'] - p1) / np.abs(p1)) < 1e-5
    J1 = one.J.cpu().numpy()
    assert np.array_equal(np.isnan(z['J']), np.isnan(J1)) and np.nanmax(np.abs(z['J'] - J1)) < 1e-5


def _stream_worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        from test_footprint import _host_scene
        scene = SyntheticScene(10, 320, 240, seed=11)
        views = list(range(10))
        host, _ = _host_scene(scene, views)
        host = host.pin()
        dev = f'cuda:{rank}'
        peers = sdist.PeerExchange(torch.device(dev))
        kw = dict(device=dev, peers=peers, num_iter=12, use_closed_form=True)
        targets = [4, 7, 2, 4]
        singles = [api.restore_from_host_sharded(host, t, views, **kw) for t in targets]
        got = list(api.restore_stream_sharded(host, targets, views, **kw))
        assert len(got) == len(targets)
        for a, b in zip(singles, got):
            assert a.n_obs == b.n_obs and a.h2d_bytes == b.h2d_bytes
            assert torch.equal(a.params.view(torch.int32), b.params.view(torch.int32))
            assert torch.equal(a.history.view(torch.int32), b.history.view(torch.int32))
            if rank == 0:
                assert torch.equal(a.J.view(torch.int32), b.J.reshape(a.J.shape).view(torch.int32))
            else:
                assert b.J is None
        if rank == 0:
            np.savez(out_path, ok=1)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs >= 2 GPUs')
def test_sharded_stream_matches_single_calls():
    """api.restore_stream_sharded (per-rank double-buffered uploads, read-back on rank 0 overlapped with the next target)
    yields, target after target and bit for bit, what api.restore_from_host_sharded returns."""
    world = 2
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, 'ok.npz')
        mp.spawn(_stream_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert os.path.exists(out)
