"""End to end through the reference-facing surface: PNG files + COLMAP model on disk -> CLI / restore_image ->
.pt and PNG outputs, compared with what the unmodified reference produced for the same scene (golden)."""
import numpy as np
import pytest
import torch

import cv2

from sucre_b200 import sfm, sucre

pytestmark = pytest.mark.gpu


def _write_golden_scene(g, root):
    for d in ('images', 'depth', 'model', 'out'):
        (root / d).mkdir()
    names = g['names'].tolist()
    cams = {}
    with open(root / 'model' / 'images.txt', 'w') as f:
        for i, name in enumerate(names):
            depth, rgb = g.inputs(i)
            cv2.imwrite(str(root / 'depth' / f'depth_{name}'), depth)
            cv2.imwrite(str(root / 'images' / name), np.ascontiguousarray(rgb[..., ::-1]))
            cam = tuple(g[f'in_cam_{i}'].tolist())
            cam_id = cams.setdefault(cam, len(cams) + 1)
            q, t = g[f'in_q_{i}'], g[f'in_t_{i}']
            f.write(' '.join([str(i + 1)] + [repr(float(x)) for x in (*q, *t)] + [str(cam_id), name]) + '\n\n')
    with open(root / 'model' / 'cameras.txt', 'w') as f:
        for cam, cam_id in cams.items():
            W, H, fx, fy, cx, cy = cam
            f.write(f'{cam_id} PINHOLE {int(W)} {int(H)} {fx!r} {fy!r} {cx!r} {cy!r}\n')
    return root


def _rel(a, b):
    return float(np.max(np.abs(a - b) / np.abs(b)))


def test_colmap_model_matches_reference_geometry(golden, tmp_path):
    g = golden('mixed8_image0004')
    root = _write_golden_scene(g, tmp_path)
    model = sfm.COLMAPModel(root / 'model', root / 'images', root / 'depth')
    for i, name in enumerate(g['names'].tolist()):
        im, a = model[name], g.geom_arrays(i)
        assert im.id == i + 1 and (im.camera.width, im.camera.height) == tuple(a['wh'])
        geom = im.geom
        for k in ('K', 'Kinv', 'R', 't', 'Ri', 'ti'):  # bit-identical to the reference's tensors
            assert np.array_equal(getattr(geom, k).numpy().reshape(a[k].shape), a[k]), (name, k)


@pytest.mark.parametrize('mode', ['closed', 'param'])
def test_cli_restores_like_the_reference(golden, tmp_path, mode, capsys):
    g = golden('tiny6_closed')
    gm = golden(f'tiny6_{mode}')
    root = _write_golden_scene(g, tmp_path)
    argv = ['--image-dir', str(root / 'images'), '--depth-dir', str(root / 'depth'), '--model-dir', str(root / 'model'),
            '--output-dir', str(root / 'out'), '--image-name', str(g['target']), '--num-iter', str(int(gm['num_iter'])),
            '--batch-size', '2', '--num-workers', '2']
    if mode == 'closed':
        argv.append('--use-closed-form')
    sucre.main(argv)
    out = capsys.readouterr().out
    assert f"Restore {g['target']}." in out and 'Total of 26797 observations.' in out and 'iter: 0024' in out
    stem = str(g['target'])[:-4]
    saved = torch.load(root / 'out' / f'{stem}.pt')
    assert set(saved) == ({'B', 'beta', 'gamma', 'J'})
    for k in ('B', 'beta', 'gamma'):
        assert saved[k].shape == (3, 1) and _rel(saved[k].numpy(), gm[k]) < 1e-4
    J = saved['J'].numpy()
    assert np.array_equal(np.isnan(J), np.isnan(gm['J'])) and np.nanmax(np.abs(J - gm['J'])) < 1e-3
    assert (root / 'out' / f'{stem}_rgb.png').exists() and (root / 'out' / f'{stem}_reconstruction.png').exists()
    assert not (root / 'out' / f'{stem}.h5').exists()  # erased without --keep-matches (sucre.py:217-219)


def test_keep_matches_and_reuse(golden, tmp_path, capsys):
    g = golden('tiny6_closed')
    root = _write_golden_scene(g, tmp_path)
    model = sfm.COLMAPModel(root / 'model', root / 'images', root / 'depth')
    image = model[str(g['target'])]
    kw = dict(colmap_model=model, output_dir=root / 'out', use_closed_form=True, num_iter=3, device='cuda')
    sucre.restore_image(image, keep_matches=True, **kw)
    assert (root / 'out' / 'image0002.h5').exists() and (root / 'out' / 'image0002.matches.npz').exists()
    first = torch.load(root / 'out' / 'image0002.pt')
    capsys.readouterr()
    sucre.restore_image(image, keep_matches=False, **kw)          # reuses the kept matches (sucre.py:185)
    assert 'Compute image0002.png matches.' not in capsys.readouterr().out
    again = torch.load(root / 'out' / 'image0002.pt')
    assert torch.equal(first['B'], again['B']) and torch.equal(first['J'].nan_to_num(), again['J'].nan_to_num())
    assert not (root / 'out' / 'image0002.matches.npz').exists()


def test_unknown_image_and_bad_camera(golden, tmp_path):
    g = golden('tiny6_closed')
    root = _write_golden_scene(g, tmp_path)
    model = sfm.COLMAPModel(root / 'model', root / 'images', root / 'depth')
    with pytest.raises(KeyError):
        model['nope.png']
    txt = (root / 'model' / 'cameras.txt').read_text().replace('PINHOLE', 'SIMPLE_RADIAL')
    (root / 'model' / 'cameras.txt').write_text(txt)
    with pytest.raises(AssertionError):
        sfm.COLMAPModel(root / 'model', root / 'images', root / 'depth')


@pytest.mark.parametrize('mode', ['closed', 'param'])
def test_image_scale_like_the_reference(golden, tmp_path, mode, capsys):
    """--image-scale 0.5: intrinsics scaled (sfm.py:193-199), depth resampled nearest, colour resampled in float
    (loader.py:158-169) and kept as float32 on the device."""
    g = golden('scaled8_closed')
    gm = golden(f'scaled8_{mode}')
    root = _write_golden_scene(g, tmp_path)
    model = sfm.COLMAPModel(root / 'model', root / 'images', root / 'depth', image_scale=0.5)
    image = model[str(g['target'])]
    assert (image.camera.width, image.camera.height) == (96, 64)
    assert np.array_equal(image.geom.K.numpy(), g.geom_arrays(g.view_index(image.name))['K'])
    sucre.restore_image(image, model, root / 'out', use_closed_form=(mode == 'closed'), num_iter=int(gm['num_iter']),
                        keep_matches=True, device='cuda')
    from sucre_b200 import loader
    mf = loader.MatchesFile(root / 'out' / 'image0004.h5', colmap_model=model)
    data = mf.load_matches(device='cuda')
    got = data.store.to_reference_layout()
    names = [n for n, k in zip(data.names, data.store.view_kept) if k]
    assert names == g['kept'].tolist()
    for key, name in zip(data.store.kept_keys, names):
        ref = g.matches(name)
        for f in ('u1', 'v1', 'u2', 'v2'):
            assert np.array_equal(got[key][f], ref[f]), (name, f)
        assert np.array_equal(got[key]['I'].view(np.uint32), ref['I'].view(np.uint32)), name   # float colour, bit-exact
        assert np.array_equal(got[key]['z'].view(np.uint32), ref['z'].view(np.uint32)), name
    saved = torch.load(root / 'out' / 'image0004.pt')
    for k in ('B', 'beta', 'gamma'):
        assert _rel(saved[k].numpy(), gm[k]) < 1e-4
    J = saved['J'].numpy()
    assert J.shape == (64, 96, 3)
    assert np.array_equal(np.isnan(J), np.isnan(gm['J'])) and np.nanmax(np.abs(J - gm['J'])) < 1e-3


def test_multi_target_cli_and_device_side_plots(golden, tmp_path):
    """--image-list: every target's files exist when the CLI returns (background writer joined); the device-side
    percentile stretch of plot_J equals the reference's numpy version to within one grey level."""
    g = golden('tiny6_closed')
    root = _write_golden_scene(g, tmp_path)
    (root / 'targets.txt').write_text('image0001.png\nimage0002.png\nimage0004.png\n')
    sucre.main(['--image-dir', str(root / 'images'), '--depth-dir', str(root / 'depth'), '--model-dir', str(root / 'model'),
                '--output-dir', str(root / 'out'), '--image-list', str(root / 'targets.txt'), '--use-closed-form',
                '--num-iter', '25'])
    for stem in ('image0001', 'image0002', 'image0004'):
        for suffix in ('.pt', '_rgb.png', '_reconstruction.png'):
            assert (root / 'out' / f'{stem}{suffix}').exists(), (stem, suffix)
    saved = torch.load(root / 'out' / 'image0002.pt')
    assert np.nanmax(np.abs(saved['J'].numpy() - g['J'])) < 1e-3
    # reference plot_J (sucre.py:84-94) in numpy on the saved J
    J = saved['J'].numpy().copy()
    valid = np.all(~np.isnan(J), axis=2)
    Jv = J[valid]
    Jv = np.clip(Jv, np.percentile(Jv, 1, axis=0), np.percentile(Jv, 99, axis=0))
    Jv = Jv - np.min(Jv, axis=0)
    Jv = Jv / np.max(Jv, axis=0)
    J[~valid] = 0.0
    J[valid] = Jv
    ref_png = np.uint8(J * 255)
    from PIL import Image as PILImage
    got_png = np.asarray(PILImage.open(root / 'out' / 'image0002_rgb.png'))
    assert got_png.shape == ref_png.shape and np.abs(got_png.astype(int) - ref_png.astype(int)).max() <= 1


def test_large_survey_decodes_only_overlapping_views(tmp_path):
    """>= 128 views in the pairing list: the target is decoded first, the conservative frustum pre-test drops the views
    it cannot reach, and only the others are decoded and uploaded; the matches are those of the unculled gather."""
    from sucre_b200 import engine, loader
    from sucre_b200.synth import SyntheticScene
    scene = SyntheticScene(144, 64, 48, seed=6)
    dirs = scene.write(tmp_path)
    model = sfm.COLMAPModel(dirs['model'], dirs['images'], dirs['depth'])
    target = model['image0000.png']                       # a corner of the 12 x 12 survey
    mf = loader.MatchesFile(tmp_path / 'image0000.h5', colmap_model=model)
    target.match_images(list(model.images.values()), mf, device='cuda')
    store = mf.store
    resident = len(model._scenes['cuda'].geom)
    assert store.stats['views_culled'] >= 60 and resident <= 85, (store.stats, resident)   # most views never decoded
    assert len(store.source_keys) == 144 and store.view_kept.sum() > 3
    full = model.scene('cuda', list(model.images.values()))               # now decode everything and gather unculled
    ordered = sorted(model.images.values(), key=lambda im: im.name)
    ref = engine.gather(full, target.id, [im.id for im in ordered], keep_src=True, cull_views=False)
    assert np.array_equal(ref.view_count, store.view_count) and np.array_equal(ref.view_kept, store.view_kept)
    assert torch.equal(ref.cells, store.cells) and torch.equal(ref.blk_view, store.blk_view)


def test_scene_residency_budget_evicts_least_recently_used_views(tmp_path):
    """The decoded views of a survey stay under COLMAPModel.scene_budget_bytes per GPU: the least recently used views a
    call does not ask for are dropped and decoded again when a later target needs them; results do not change."""
    from sucre_b200 import engine
    from sucre_b200.synth import SyntheticScene
    scene = SyntheticScene(12, 64, 48, seed=3)
    dirs = scene.write(tmp_path)
    model = sfm.COLMAPModel(dirs['model'], dirs['images'], dirs['depth'])
    ims = list(model.images.values())
    per_view = 5 * 64 * 48
    model.scene_budget_bytes = 6 * per_view
    sc = model.scene('cuda', ims[:5])
    assert len(sc.geom) == 5
    sc = model.scene('cuda', ims[3:9])                       # 6 wanted, 2 already there: the other 3 must go
    assert len(sc.geom) == 6 and set(sc.geom) == {im.id for im in ims[3:9]}
    sc = model.scene('cuda', ims[:2])                        # budget allows 6: the two oldest of the residents leave
    assert len(sc.geom) == 6 and {ims[0].id, ims[1].id} <= set(sc.geom)
    a = engine.gather(sc, ims[0].id, [ims[0].id, ims[1].id])
    model.scene_budget_bytes = None
    model.drop_scene()
    b = engine.gather(model.scene('cuda', ims), ims[0].id, [ims[0].id, ims[1].id])
    assert a.n_obs == b.n_obs and torch.equal(a.cells, b.cells)
