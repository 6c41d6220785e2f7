"""Parity of the CUDA gather (through the C ABI) with the oracle and with the reference's own outputs."""
import numpy as np
import pytest
import torch

from sucre_b200 import engine
from sucre_b200.synth import SyntheticScene

import helpers

pytestmark = pytest.mark.gpu

FULL = ['tiny6_closed', 'mixed8_image0004', 'mixed8_image0002']


@pytest.mark.parametrize('case', FULL)
def test_gather_vs_reference_golden(golden, case):
    """Bit-exact against arrays captured from the unmodified reference (indices, z, I; kept views; order)."""
    g = golden(case)
    ds, _ = helpers.golden_device_scene(g)
    names = g['names'].tolist()
    order = sorted((names[i] for i in g.pairing_list()))
    store = engine.gather(ds, str(g['target']), order, min_cover=float(g['min_cover']), keep_src=True)
    assert store.kept_keys == g['kept'].tolist()
    got = store.to_reference_layout()
    total = 0
    for name in g['kept'].tolist():
        ref = g.matches(name)
        for f in ('u1', 'v1', 'u2', 'v2'):
            assert np.array_equal(got[name][f], ref[f]), (name, f)
        assert np.array_equal(got[name]['z'].view(np.uint32), ref['z'].view(np.uint32)), (name, 'z')
        assert np.array_equal(got[name]['I'].view(np.uint32), ref['I'].view(np.uint32)), (name, 'I')
        total += len(ref['u1'])
    assert store.n_obs == total == len(store)


def test_gather_vs_oracle_config1():
    """BASELINE.json configs[0] shape (20 views 640x480): every pixel-view against the oracle."""
    scene = SyntheticScene(20, 640, 480, seed=0)
    ds, host = helpers.build_device_scene(scene, range(20), render_device='cuda')
    kept, stats = helpers.oracle_gather(host, 8, list(range(20)))
    store = engine.gather(ds, 8, list(range(20)), keep_src=True)
    assert store.view_count.tolist() == [stats[i][0] for i in range(20)]
    bad = helpers.compare_store_with_oracle(store, kept)
    assert bad['idx'] == 0 and bad['z'] == 0 and bad['I'] == 0, bad
    assert bad['n'] == store.n_obs > 3_000_000


def test_gather_ragged_and_degenerate():
    """Edge cases: image size not a multiple of the 32-pixel tile, a target that is not in the pairing list,
    a min_cover that drops every view, views with all-invalid depth."""
    scene = SyntheticScene(5, 75, 41, seed=5)  # 3075 pixels = 96 tiles + 3 pixels
    ds, host = helpers.build_device_scene(scene, range(5))
    # target excluded from its own pairing list: some valid pixels end up unobserved
    src = [0, 1, 3, 4]
    kept, stats = helpers.oracle_gather(host, 2, src)
    store = engine.gather(ds, 2, src, keep_src=True)
    assert helpers.compare_store_with_oracle(store, kept) == dict(idx=0, z=0, I=0, n=store.n_obs)
    # min_cover = 1.0 can never be exceeded -> empty store
    empty = engine.gather(ds, 2, src, min_cover=1.0)
    assert empty.n_obs == 0 and empty.n_blocks == 0 and not empty.view_kept.any()
    assert empty.view_count.tolist() == [stats[i][0] for i in src]
    # a view whose depth map is all zero never matches and is dropped
    geom = ds.geom[1]
    ds.add_view('blank', geom, torch.zeros((geom.height, geom.width), dtype=torch.uint16), ds.rgb[1])
    store2 = engine.gather(ds, 2, [0, 'blank', 1])
    assert store2.view_count[1] == 0 and not store2.view_kept[1] and store2.view_kept[0] and store2.view_kept[2]


def test_gather_api_errors():
    scene = SyntheticScene(2, 64, 48, seed=1)
    ds, _ = helpers.build_device_scene(scene, range(2))
    with pytest.raises(engine._lib.SucreError):
        engine.gather(ds, 0, [])
    with pytest.raises(engine._lib.SucreError):
        engine.DeviceScene('cpu')
    L = engine._lib.lib()
    assert L.sucre_gather_match(0, 0, 1, 0, 1, 0, 0, 0) != 0 and b'null' in L.sucre_last_error()


def test_view_culling_changes_nothing():
    """The conservative frustum pre-test skips views that cannot overlap; the store is the same with and without it."""
    scene = SyntheticScene(49, 96, 64, seed=8)          # 7x7 grid: a corner target sees only its neighbourhood
    ds, host = helpers.build_device_scene(scene, range(49))
    keys = list(range(49))
    a = engine.gather(ds, 0, keys, keep_src=True, cull_views=True)
    b = engine.gather(ds, 0, keys, keep_src=True, cull_views=False)
    assert a.stats['views_culled'] >= 8 and 'views_culled' not in b.stats
    assert np.array_equal(a.view_count, b.view_count) and np.array_equal(a.view_kept, b.view_kept) and a.n_obs == b.n_obs
    assert torch.equal(a.cells, b.cells) and torch.equal(a.blk_view, b.blk_view) and torch.equal(a.blk_mask, b.blk_mask)
    keep = ds.possibly_overlapping(0, keys)
    assert b.view_count[~keep].sum() == 0              # conservative: nothing with matches was culled
    kept, stats = helpers.oracle_gather(host, 0, keys)
    assert helpers.compare_store_with_oracle(a, kept) == dict(idx=0, z=0, I=0, n=a.n_obs)


@pytest.mark.parametrize('shape', [(7, 150, 101), (5, 64, 48), (9, 320, 240), (300, 64, 48)])   # 300 views: two passes of the mask staging
def test_slot_permutation_changes_only_the_layout(shape):
    """sucre_gather_permute deals the pixels of every 32-tile group to the slots by observation count: the store then
    holds the same records (per view, in target order), every pixel exactly once in `pix`, columns of non-increasing
    height inside a group, fewer rows — and the fit gives the same J and parameters up to fp32 summation order."""
    V, W, H = shape
    scene = SyntheticScene(V, W, H, seed=21)
    ds, host = helpers.build_device_scene(scene, range(V))
    keys = list(range(V))
    t = V // 2
    a = engine.gather(ds, t, keys, keep_src=True, permute=True, cull_views=False)
    b = engine.gather(ds, t, keys, keep_src=True, permute=False, cull_views=False)
    assert a.pix is not None and b.pix is None and a.n_obs == b.n_obs and a.n_rows <= b.n_rows
    assert np.array_equal(a.view_count, b.view_count) and np.array_equal(a.view_kept, b.view_kept)
    pix = a.pix.cpu().numpy()
    assert pix.shape == (a.n_tiles * 32,) and np.array_equal(np.sort(pix[pix >= 0]), np.arange(W * H))
    la, lb = a.to_reference_layout(), b.to_reference_layout()
    assert la.keys() == lb.keys()
    for key in la:
        for f in ('u1', 'v1', 'u2', 'v2', 'z', 'I'):
            assert np.array_equal(la[key][f], lb[key][f]), (key, f)
    # column heights: non-increasing over the slots of a group (count over ALL listed views decides; here all are kept)
    if a.view_kept.all():
        _, pixel, _ = a.record_index()
        cnt = torch.bincount(pixel, minlength=W * H).cpu().numpy()
        per_slot = np.where(pix >= 0, cnt[np.maximum(pix, 0)], -1)
        for g in range(0, len(per_slot), 1024):
            grp = per_slot[g:g + 1024]
            assert (np.diff(grp) <= 0).all(), g
    kept, _ = helpers.oracle_gather(host, t, keys)
    assert helpers.compare_store_with_oracle(a, kept) == dict(idx=0, z=0, I=0, n=a.n_obs)
    for closed in (True, False):
        out = []
        for store in (a, b):
            J0 = None if closed else ds.rgb_float(t)
            st = engine.FitState.initial(ds.device, J0=J0)
            hist = engine.fit(store, st, 30)
            J = engine.closed_form_J(store, st.params, st.J) if closed else st.J
            out.append((hist.cpu().numpy(), J.cpu().numpy()))
        assert np.allclose(out[0][0], out[1][0], rtol=2e-5, atol=1e-7)
        assert np.array_equal(np.isnan(out[0][1]), np.isnan(out[1][1]))
        assert np.nanmax(np.abs(out[0][1] - out[1][1])) < 2e-5


def test_degenerate_inputs_follow_the_reference_rules():
    """NaN pose -> every comparison false -> no match (the reference's .long() of NaN is INT64_MIN, rejected);
    a target without any valid depth -> nothing to restore; oversized images are refused (int16 indices)."""
    scene = SyntheticScene(3, 64, 48, seed=3)
    ds, host = helpers.build_device_scene(scene, range(3))
    g = ds.geom[1]
    bad = engine.ViewGeom(K=g.K, Kinv=g.Kinv, R=g.R, t=g.t, Ri=g.Ri.clone(), ti=g.ti.clone(), width=g.width, height=g.height)
    bad.Ri[0, 0] = float('nan')
    ds.add_view('nan', bad, ds.depth[1], ds.rgb[1])
    store = engine.gather(ds, 0, [0, 'nan', 2])
    assert store.view_count[1] == 0 and not store.view_kept[1] and store.view_kept[0]
    ds.add_view('blind', ds.geom[0], torch.zeros((48, 64), dtype=torch.uint16), ds.rgb[0])
    empty = engine.gather(ds, 'blind', [0, 1, 2])
    assert empty.n_obs == 0
    from sucre_b200 import api
    with pytest.raises(engine._lib.SucreError):
        api.restore_resident(ds, 'blind', [0, 1, 2], num_iter=2)
    huge = engine.ViewGeom(K=g.K, Kinv=g.Kinv, R=g.R, t=g.t, Ri=g.Ri, ti=g.ti, width=40000, height=1)
    ds.geom['huge'], ds.depth['huge'], ds.rgb['huge'] = huge, torch.zeros(40000, dtype=torch.int16, device='cuda'), None
    with pytest.raises(engine._lib.SucreError, match='32767'):
        engine.gather(ds, 'huge', [0])
