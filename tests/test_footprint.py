"""CPU: the footprint upload plan (api.upload_plan / engine.DeviceScene.footprints) is conservative — the oracle's
gather gives the same observations, bit for bit, when everything outside the rectangles is erased from the source
views (which is what the device sees after a footprint upload: zero = invalid depth)."""
import numpy as np
import pytest
import torch

import helpers
from oracle import oracle
from sucre_b200 import api
from sucre_b200.engine import DeviceScene, ViewGeom
from sucre_b200.synth import SyntheticScene


def _host_scene(scene, views):
    geoms, depth, rgb, og = [], [], [], []
    for i in views:
        K, R, t, W, H = helpers.reference_pose(scene, i)
        g = ViewGeom.from_pose(K, R, t, W, H)
        d, c = scene.render(i)
        geoms.append(g)
        depth.append(d)
        rgb.append(c)
        og.append(oracle.view_geom(g.K, g.R, g.t, W, H, Kinv=g.Kinv, Ri=g.Ri, ti=g.ti))
    return api.HostScene(geoms, torch.stack(depth), torch.stack(rgb)), og


def _erase_outside(a: np.ndarray, rect) -> np.ndarray:
    x0, y0, x1, y1 = (int(v) for v in rect)
    out = np.zeros_like(a)
    out[y0:y1, x0:x1] = a[y0:y1, x0:x1]
    return out


@pytest.mark.parametrize('n_views,width,height,seed,target,mode', [
    (12, 160, 120, 4, 5, 'footprint'), (12, 160, 120, 4, 0, 'rows'), (20, 200, 136, 1, 11, 'footprint'),
    (9, 96, 128, 7, 8, 'footprint')])
def test_erasing_everything_outside_the_footprints_changes_nothing(n_views, width, height, seed, target, mode):
    scene = SyntheticScene(n_views, width, height, seed=seed)
    views = list(range(n_views))
    host, og = _host_scene(scene, views)
    needed, rects = api.upload_plan(host, target, views, mode)
    assert needed == views and rects.shape == (n_views, 4) and rects.dtype == np.int32
    assert tuple(rects[target]) == (0, 0, width, height)        # the target itself travels whole
    area = (rects[:, 2] - rects[:, 0]).astype(np.int64) * (rects[:, 3] - rects[:, 1])
    assert area.sum() < (0.95 if mode == 'rows' else 0.8) * n_views * width * height   # the plan saves something here
    assert api.h2d_bytes(host, target, views, mode) == 5 * area.sum()
    assert api.h2d_bytes(host, target, views, 'full') == 5 * n_views * width * height
    if mode == 'rows':
        assert all(r[0] == 0 and r[2] == width for r in rects if r[3] > r[1])

    dT = host.depth[target].numpy().view(np.uint16)
    full = [(i, host.depth[i].numpy().view(np.uint16), host.rgb[i].numpy(), og[i]) for i in views]
    part = [(i, _erase_outside(d, rects[i]), _erase_outside(c, rects[i]), g) for i, d, c, g in full]
    kept_f, stats_f = oracle.gather(dT, og[target], full)
    kept_p, stats_p = oracle.gather(dT, og[target], part)
    assert [k for k, _ in kept_f] == [k for k, _ in kept_p] and len(kept_f) >= 3
    assert {k: v[0] for k, v in stats_f.items()} == {k: v[0] for k, v in stats_p.items()}   # matches per view
    for (_, a), (_, b) in zip(kept_f, kept_p):
        for f in ('u1', 'v1', 'u2', 'v2', 'd', 'z', 'I', 'cP'):
            assert np.array_equal(a[f].view(np.uint8), b[f].view(np.uint8)), f
    # every match lies inside its view's rectangle; views with an empty rectangle had none
    for key, obs in kept_f:
        x0, y0, x1, y1 = rects[key]
        assert (obs['u2'] >= x0).all() and (obs['u2'] < x1).all() and (obs['v2'] >= y0).all() and (obs['v2'] < y1).all()
    for i in views:
        if area[i] == 0:
            assert stats_f[i][1] == 0                            # not even an in-bounds forward projection


def test_every_forward_projection_lands_inside_the_footprint():
    """Stronger than the matches: no in-bounds forward projection (float64 restatement of sfm.py:90-107, 116-117) of
    any valid target pixel falls outside the rectangle, with a pixel to spare."""
    scene = SyntheticScene(16, 192, 128, seed=3)
    views = list(range(16))
    host, _ = _host_scene(scene, views)
    target = 6
    g = host.geoms[target]
    rects = DeviceScene.footprints(g, api.host_depth_range(host.depth[target]), host.geoms)
    d = host.depth[target].numpy().view(np.uint16).astype(np.float64) / 1000.0
    v, u = np.nonzero(d > 0)
    X = np.stack([(u + 0.5) * d[v, u], (v + 0.5) * d[v, u], d[v, u]])
    wP = g.R.double().numpy() @ (g.Kinv.double().numpy() @ X) + g.t.double().numpy()
    partial = 0
    for s, gs in enumerate(host.geoms):
        p = gs.K.double().numpy() @ (gs.Ri.double().numpy() @ wP + gs.ti.double().numpy())
        px, py = p[0] / p[2], p[1] / p[2]
        inb = (px > -1) & (px < gs.width) & (py > -1) & (py < gs.height)
        x0, y0, x1, y1 = rects[s]
        if inb.any():
            assert px[inb].min() >= x0 + 1 or x0 == 0
            assert py[inb].min() >= y0 + 1 or y0 == 0
            assert px[inb].max() <= x1 - 1 or x1 == gs.width
            assert py[inb].max() <= y1 - 1 or y1 == gs.height
        partial += (x1 - x0) * (y1 - y0) < gs.width * gs.height
    assert partial >= 8


def test_footprints_degenerate_inputs():
    scene = SyntheticScene(6, 64, 48, seed=2)
    host, _ = _host_scene(scene, range(6))
    g = host.geoms[0]
    whole = np.array([[0, 0, 64, 48]] * 6, dtype=np.int32)
    # a target without any valid depth: nothing can be concluded, whole views
    assert np.array_equal(DeviceScene.footprints(g, (65.536, 0.0), host.geoms), whole)
    assert api.host_depth_range(torch.zeros((48, 64), dtype=torch.int16)) == (65.536, 0.0)
    # a camera behind the target's slab is not bounded by the projection argument: whole view
    behind = ViewGeom.from_pose(g.K, g.R, g.t + g.R @ torch.tensor([[0.0], [0.0], [3.0]]), 64, 48)
    r = DeviceScene.footprints(g, (1.5, 2.5), [behind, g])
    assert tuple(r[0]) == (0, 0, 64, 48) and tuple(r[1]) == (0, 0, 64, 48)
    # far away along the camera's own x axis (same image plane): empty
    far = ViewGeom.from_pose(g.K, g.R, g.t + g.R @ torch.tensor([[50.0], [0.0], [0.0]]), 64, 48)
    assert tuple(DeviceScene.footprints(g, (1.5, 2.5), [far])[0]) == (0, 0, 0, 0)
    assert DeviceScene.footprints(g, (1.5, 2.5), []).shape == (0, 4)
    with pytest.raises(ValueError):
        api.upload_plan(host, 0, [1, 2], 'everything')
    needed, rects = api.upload_plan(host, 2, [4, 1], 'full')
    assert needed == [1, 2, 4] and np.array_equal(rects, whole[:3])
    depth = host.depth[3]
    lo, hi = api.host_depth_range(depth)
    a = depth.numpy().view(np.uint16)
    assert lo == a[a > 0].min() / 1000.0 and hi == a.max() / 1000.0


def test_view_level_pre_test_agrees_with_the_oracle_and_the_footprints():
    """DeviceScene.possibly_overlapping (the view-level cull of engine.gather) on the host only: a culled view has
    no in-bounds forward projection in the oracle, and its footprint rectangle is empty."""
    scene = SyntheticScene(64, 96, 64, seed=5)
    views = list(range(64))
    host, og = _host_scene(scene, views)
    target = 0
    ds = object.__new__(DeviceScene)                 # no device: geometry and the cached depth range are all it needs
    ds.geom = dict(enumerate(host.geoms))
    ds._ranges = {target: api.host_depth_range(host.depth[target])}
    ds._cull_tables = {}
    keep = ds.possibly_overlapping(target, views)
    assert keep.dtype == bool and keep[target] and 8 <= (~keep).sum() < 64
    assert np.array_equal(keep, ds.possibly_overlapping(target, views, source_geoms=host.geoms))   # cached tables
    rects = DeviceScene.footprints(host.geoms[target], ds._ranges[target], host.geoms)
    dT = host.depth[target].numpy().view(np.uint16)
    for s in views:
        if not keep[s]:
            _, n, n_in = oracle.match_pair(dT, og[target], host.depth[s].numpy().view(np.uint16), og[s])
            assert n == 0 and n_in == 0
            assert rects[s, 2] - rects[s, 0] <= 3 or rects[s, 3] - rects[s, 1] <= 3   # at most the 2-pixel margin
