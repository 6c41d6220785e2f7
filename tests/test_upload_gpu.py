"""GPU: restore_from_host gives the same result bit for bit whether the source views are uploaded whole or only
their footprint rectangles (sucre_scene_upload), and both agree with a scene that was resident from the start."""
import numpy as np
import pytest
import torch

from sucre_b200 import api, engine
from test_footprint import _host_scene
from sucre_b200.synth import SyntheticScene

pytestmark = pytest.mark.gpu


def _same(a: api.RestoreResult, b: api.RestoreResult):
    assert a.n_obs == b.n_obs and np.array_equal(a.view_kept, b.view_kept)
    for f in ('J', 'params', 'history'):
        x, y = getattr(a, f).cpu().numpy(), getattr(b, f).cpu().numpy()
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32)), f


@pytest.mark.parametrize('closed', [True, False])
def test_footprint_upload_is_bit_identical_to_full_upload(closed):
    scene = SyntheticScene(20, 200, 136, seed=1)
    views = list(range(20))
    host, _ = _host_scene(scene, views)
    host = host.pin()
    kw = dict(device='cuda:0', min_cover=1e-6, use_closed_form=closed, num_iter=12, lr=0.05)
    for target, sources in ((11, views), (3, [0, 1, 2, 3, 4, 7, 8, 9, 12, 17])):
        full = api.restore_from_host(host, target, sources, upload='full', **kw)
        assert full.h2d_bytes == api.h2d_bytes(host, target, sources, 'full') and full.n_obs > 10000
        for mode in ('footprint', 'rows'):
            part = api.restore_from_host(host, target, sources, upload=mode, **kw)
            _same(full, part)
            assert part.h2d_bytes == api.h2d_bytes(host, target, sources, mode) < full.h2d_bytes
        resident = engine.DeviceScene('cuda:0')
        for i in views:
            resident.add_view(i, host.geoms[i], host.depth[i], host.rgb[i])
        _same(full, api.restore_resident(resident, target, sources, **{k: v for k, v in kw.items() if k != 'device'}))


def test_stream_pipeline_matches_single_calls_and_never_reads_outside_the_rectangles():
    """api.restore_stream (double-buffered uploads on a copy stream, read-back overlapped) yields, target after target,
    exactly what restore_from_host returns.  Its scene buffers are reused without being cleared, so whatever earlier
    targets left outside the current target's rectangles must never be read: the results stay bit-identical."""
    scene = SyntheticScene(20, 200, 136, seed=1)
    views = list(range(20))
    host, _ = _host_scene(scene, views)
    host = host.pin()
    kw = dict(min_cover=1e-6, use_closed_form=True, num_iter=8, lr=0.05)
    targets = [11, 3, 18, 0, 11, 7]
    singles = [api.restore_from_host(host, t, views, device='cuda:0', upload='full', **kw) for t in targets]
    for mode in ('footprint', 'full'):
        got = list(api.restore_stream(host, targets, views, device='cuda:0', upload=mode, **kw))
        assert len(got) == len(targets)
        for a, b, t in zip(singles, got, targets):
            _same(a, b)
            assert b.h2d_bytes == api.h2d_bytes(host, t, views, mode)
    out = [torch.empty((136, 200, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
    for k, r in enumerate(api.restore_stream(host, targets[:3], views, device='cuda:0', out_J=out, **kw)):
        assert r.J is out[k % 2]
        _same(singles[k], r)


def test_scene_upload_copies_exactly_the_rectangles():
    H, W, n = 37, 53, 5
    g = torch.Generator().manual_seed(0)
    src = torch.randint(1, 255, (n, H, W, 3), dtype=torch.uint8, generator=g).pin_memory()
    depth = torch.randint(1, 30000, (n, H, W), dtype=torch.int16, generator=g).pin_memory()
    rects = np.array([[0, 0, W, H], [5, 7, 20, 30], [0, 3, W, 9], [10, 10, 10, 20], [W - 1, H - 1, W, H]], dtype=np.int32)
    order = [4, 2, 0, 1, 3]
    geoms = [engine.ViewGeom.from_pose(torch.eye(3), torch.eye(3), torch.zeros(3, 1), W, H)] * n
    scene = engine.DeviceScene('cuda:0')
    planes = scene.allocate_views(list('abcde'), geoms, depth.dtype)
    planes[1].zero_()  # colour planes come uninitialised
    copied = scene.upload_rects(planes, depth, src, order, rects)
    area = (rects[:, 2] - rects[:, 0]) * (rects[:, 3] - rects[:, 1])
    assert copied == 5 * int(area.sum())
    for i, key in enumerate('abcde'):
        x0, y0, x1, y1 = rects[i]
        want_c, want_d = torch.zeros((H, W, 3), dtype=torch.uint8), torch.zeros((H, W), dtype=torch.int16)
        want_c[y0:y1, x0:x1] = src[order[i], y0:y1, x0:x1]
        want_d[y0:y1, x0:x1] = depth[order[i], y0:y1, x0:x1]
        assert torch.equal(scene.rgb[key].cpu(), want_c) and torch.equal(scene.depth[key].cpu(), want_d)
    bad = rects.copy()
    bad[1] = [5, 7, W + 1, 30]
    with pytest.raises(engine._lib.SucreError, match='outside'):
        scene.upload_rects(planes, depth, src, order, bad)
