"""CPU-only tests of the host side that mirrors the reference's interface: COLMAP reader (text + binary), intrinsics
scaling, file naming, loaders, CLI flags and defaults, target/pairing selection."""
import numpy as np
import pytest
import torch

from sucre_b200 import loader, sfm, sucre
from sucre_b200.synth import SyntheticScene


@pytest.fixture(scope='module')
def scene_dir(tmp_path_factory):
    root = tmp_path_factory.mktemp('scene')
    scene = SyntheticScene(5, 64, 48, seed=2, alt_size=(48, 40), alt_every=2)
    dirs = scene.write(root, binary_model=True)
    return scene, root, dirs


def test_text_and_binary_models_agree(scene_dir, tmp_path):
    scene, root, dirs = scene_dir
    cams_b, imgs_b = sfm.read_colmap_model(dirs['model'])          # binary wins when both exist
    txt = tmp_path / 'txt'
    txt.mkdir()
    for f in ('cameras.txt', 'images.txt'):
        (txt / f).write_text((dirs['model'] / f).read_text())
    cams_t, imgs_t = sfm.read_colmap_model(txt)
    assert cams_b == cams_t
    assert list(imgs_b) == list(imgs_t) == [1, 2, 3, 4, 5]
    for k in imgs_b:
        assert imgs_b[k]['name'] == imgs_t[k]['name'] and imgs_b[k]['camera_id'] == imgs_t[k]['camera_id']
        assert np.allclose(imgs_b[k]['qvec'], imgs_t[k]['qvec'], rtol=0, atol=0)
        assert np.allclose(imgs_b[k]['tvec'], imgs_t[k]['tvec'], rtol=0, atol=0)
    with pytest.raises(FileNotFoundError):
        sfm.read_colmap_model(tmp_path / 'missing')


def test_colmap_model_surface(scene_dir):
    scene, root, dirs = scene_dir
    model = sfm.COLMAPModel(dirs['model'], dirs['images'], dirs['depth'])
    assert repr(model) == 'COLMAPModel(5 images)' and set(model.cameras) == {1, 2}
    im = model['image0003.png']
    assert im.id == 4 and im.name == 'image0003.png' and im.camera.id == 2 and (im.camera.width, im.camera.height) == (48, 40)
    assert im.depth_map_path == dirs['depth'] / 'depth_image0003.png'      # sfm.py:214
    assert im.pose.R.dtype == torch.float32 and im.pose.t.shape == (3, 1)
    # the stored pose is cam->world: its inverse maps the camera centre to the origin
    C = torch.tensor(scene.C[3], dtype=torch.float32).view(3, 1)
    assert torch.allclose(im.pose.inverse().transform(C), torch.zeros(3, 1), atol=1e-5)
    assert torch.allclose(im.pose.transform(torch.zeros(3, 1)), C, atol=1e-5)
    g = im.geom
    assert torch.equal(g.Kinv, im.camera.K.inverse()) and torch.equal(g.Ri, im.pose.R.T) and torch.equal(g.ti, -im.pose.R.T @ im.pose.t)
    with pytest.raises(KeyError):
        model['nope.png']


def test_image_scale_scales_intrinsics_like_the_reference(scene_dir):
    scene, root, dirs = scene_dir
    model = sfm.COLMAPModel(dirs['model'], dirs['images'], dirs['depth'], image_scale=0.3)
    cam = model.cameras[1]
    W, H, fx, fy, cx, cy = scene.cams[0]
    w, h = int(W * 0.3), int(H * 0.3)                                      # sfm.py:193-199
    assert (cam.width, cam.height) == (w, h)
    K = torch.tensor([[fx * (w / W), 0, cx * (w / W)], [0, fy * (h / H), cy * (h / H)], [0, 0, 1]], dtype=torch.float32)
    assert torch.equal(cam.K, K)
    im = model['image0000.png']
    rgb, depth = im.get_rgb_device_form(), im.get_depth_u16()
    assert rgb.dtype == torch.float32 and rgb.shape == (h, w, 3) and depth.dtype == torch.uint16 and depth.shape == (h, w)
    assert torch.equal(rgb, im.get_rgb())                                   # same resampling as loader.py:157-163
    assert torch.equal(depth.to(torch.int32).float() / 1000, im.get_depth_map())


def test_loaders_match_the_reference_formulas(scene_dir):
    scene, root, dirs = scene_dir
    depth_u16, rgb_u8 = scene.render(0)
    p_rgb, p_depth = dirs['images'] / 'image0000.png', dirs['depth'] / 'depth_image0000.png'
    assert torch.equal(loader.load_depth_u16(p_depth, 64, 48), depth_u16)
    assert torch.equal(loader.load_rgb_device_form(p_rgb, 64, 48), rgb_u8)
    assert torch.equal(loader.load_depth_map(p_depth, 64, 48), torch.tensor(depth_u16.numpy() / 1000, dtype=torch.float32))
    assert torch.equal(loader.load_rgb(p_rgb, 64, 48), torch.tensor(rgb_u8.numpy() / 255, dtype=torch.float32))
    with pytest.raises(FileNotFoundError):
        loader.load_depth_u16(dirs['depth'] / 'nope.png', 64, 48)
    with pytest.raises(ValueError):
        loader.load_depth_u16(p_rgb, 64, 48)                                # not a single-channel map
    import cv2
    d8 = (depth_u16.numpy() >> 8).astype('uint8')                            # an 8-bit depth map: the reference divides it by 1000 all the same
    cv2.imwrite(str(root / 'depth8.png'), d8)
    assert torch.equal(loader.load_depth_u16(root / 'depth8.png', 64, 48), torch.from_numpy(d8.astype('uint16')))
    assert torch.equal((loader.load_depth_u16(root / 'depth8.png', 64, 48).to(torch.int32).to(torch.float64) / 1000).to(torch.float32),
                       loader.load_depth_map(root / 'depth8.png', 64, 48))


def test_cli_flags_and_defaults_are_the_reference_ones():
    p = sucre.build_parser()
    req = ['--image-dir', 'i', '--depth-dir', 'd', '--model-dir', 'm', '--output-dir', 'o']
    a = p.parse_args(req + ['--image-name', 'x.png'])
    assert (a.min_cover, a.image_scale, a.learning_rate, a.num_iter, a.batch_size, a.num_workers, a.device) == \
        (1e-6, 1.0, 0.05, 200, 5, 0, 'cuda')                                # sucre.py:282-305
    assert not (a.light_model or a.use_closed_form or a.force_compute_matches or a.keep_matches)
    assert a.save_interval is None and a.params_path is None and a.filter_images_path is None
    assert p.parse_args(req + ['--image-ids', '3', '9']).image_ids == [3, 9]
    for bad in ([], ['--image-name', 'a', '--image-ids', '1', '2'], ['--image-list', 'l.txt', '--image-name', 'a']):
        with pytest.raises(SystemExit):
            p.parse_args(req + bad)                                         # exactly one target selector, sucre.py:272-277
    with pytest.raises(SystemExit):
        p.parse_args(['--image-name', 'x.png'])                             # the four directories are required


def test_target_and_pairing_selection(scene_dir, tmp_path, monkeypatch):
    """parse_args: --image-ids is [min, max) with missing ids skipped; --image-list reads names; filtered images leave
    the pairing list but can still be targets (sucre.py:228-239)."""
    scene, root, dirs = scene_dir
    calls = []
    monkeypatch.setattr(sucre, 'restore_image', lambda **kw: calls.append(kw))
    base = ['--image-dir', str(dirs['images']), '--depth-dir', str(dirs['depth']), '--model-dir', str(dirs['model']),
            '--output-dir', str(tmp_path / 'out')]
    sucre.main(base + ['--image-ids', '4', '9'])
    assert [c['image'].id for c in calls] == [4, 5] and (tmp_path / 'out').is_dir()
    assert [im.id for im in calls[0]['image_list']] == [1, 2, 3, 4, 5] and calls[0]['batch_size'] == 5
    calls.clear()
    (tmp_path / 'targets.txt').write_text('image0001.png\nimage0004.png\n')
    (tmp_path / 'skip.txt').write_text('image0001.png\nimage0002.png\n')
    sucre.main(base + ['--image-list', str(tmp_path / 'targets.txt'), '--filter-images-path', str(tmp_path / 'skip.txt'),
                       '--use-closed-form', '--min-cover', '0.01', '--num-iter', '7'])
    assert [c['image'].name for c in calls] == ['image0001.png', 'image0004.png']
    assert [im.name for im in calls[0]['image_list']] == ['image0000.png', 'image0003.png', 'image0004.png']
    assert calls[0]['use_closed_form'] and calls[0]['min_cover'] == 0.01 and calls[0]['num_iter'] == 7


def test_cli_shards_targets_over_torchrun_ranks(scene_dir, tmp_path, monkeypatch):
    """Under torchrun (WORLD_SIZE / RANK / LOCAL_RANK) every rank takes a contiguous share of the targets on its own
    GPU; the shares cover the target list exactly once and the pairing list is untouched (sucre.py:243-261 is a loop
    over independent targets)."""
    scene, root, dirs = scene_dir
    base = ['--image-dir', str(dirs['images']), '--depth-dir', str(dirs['depth']), '--model-dir', str(dirs['model']),
            '--output-dir', str(tmp_path / 'out'), '--image-ids', '1', '6']
    seen = []
    for rank in range(3):
        calls = []
        monkeypatch.setattr(sucre, 'restore_image', lambda **kw: calls.append(kw))
        for k, v in (('WORLD_SIZE', '3'), ('RANK', str(rank)), ('LOCAL_RANK', str(rank))):
            monkeypatch.setenv(k, v)
        sucre.main(base)
        assert all(c['device'] == f'cuda:{rank}' for c in calls)
        assert all([im.id for im in c['image_list']] == [1, 2, 3, 4, 5] for c in calls)
        seen.append([c['image'].id for c in calls])
    assert sum(seen, []) == [1, 2, 3, 4, 5] and max(map(len, seen)) - min(map(len, seen)) <= 1
    from sucre_b200.dist import shard_targets
    assert [shard_targets(list(range(10)), r, 4, contiguous=True) for r in range(4)] == [[0, 1], [2, 3, 4], [5, 6], [7, 8, 9]]


def test_no_cpu_path():
    from sucre_b200 import engine
    with pytest.raises(engine._lib.SucreError):
        engine.DeviceScene('cpu')


def test_log_formatter_matches_numpy_printing():
    """The per-iteration log line (sucre.py:149-152) prints B, beta, gamma with np.printoptions(precision=4)."""
    rng = np.random.default_rng(0)
    for _ in range(3000):
        x = (rng.normal(size=3) * 10.0 ** rng.integers(-6, 4)).astype(np.float32)
        if rng.random() < 0.2:
            x = np.round(x, 1)
        if rng.random() < 0.05:
            x[rng.integers(3)] = 0
        with np.printoptions(precision=4):
            assert sucre._fmt(x) == str(x), x
    assert sucre._fmt([0.1198, 0.161, 0.1593]) == '[0.1198 0.161  0.1593]'
    assert 'nan' in sucre._fmt([float('nan'), 0.1, 0.2])


def test_async_writer_runs_jobs_and_reraises(tmp_path):
    done = []
    with loader.AsyncWriter(2) as w:
        for i in range(8):
            w.submit(lambda k: done.append(k), i)
    assert sorted(done) == list(range(8))

    def boom():
        raise OSError('disk full')
    w = loader.AsyncWriter(1)
    w.submit(boom)
    with pytest.raises(OSError):
        w.close()


def test_store_slot_and_pixel_views_of_a_permuted_store():
    """Host-side bookkeeping of a store whose slots were dealt by observation count (engine.ObservationStore.pix): the
    shapes callers see, the struct the C ABI gets, and the image <-> slot copies of the J-parameter mode."""
    import numpy as np
    import torch
    from sucre_b200 import _lib, engine
    W, H = 10, 7                                   # 70 pixels -> 3 tiles, 96 slots, 26 of them without a pixel
    n_tiles = 3
    rng = np.random.default_rng(0)
    pix = np.full(n_tiles * 32, -1, np.int32)
    pix[rng.permutation(96)[:70]] = rng.permutation(70)
    z64 = torch.zeros(n_tiles + 1, dtype=torch.int64)

    def store(pix_t, band=None):
        return engine.ObservationStore(width=W, height=H, source_keys=(0,), view_count=np.zeros(1, np.int64), view_kept=np.ones(1, bool),
                                       n_obs=0, n_blocks=0, n_rows=0, cells=torch.zeros((1, 32, 2)), row_off=z64, blk_off=z64, rec_off=z64,
                                       blk_mask=torch.zeros(0, dtype=torch.int32), blk_view=torch.zeros(0, dtype=torch.int32), cell_src=None,
                                       band=band, pix=pix_t)

    s = store(torch.from_numpy(pix))
    assert s.n_slots == 96 and s.J_shape == (H, W, 3) and s.slot_J_shape == (96, 3) and s.needs_slot_copy and not s.is_band
    cs = s.c_struct()
    assert cs.pixels == W * H and cs.pix == s.pix.data_ptr() and cs.n_tiles == n_tiles
    img = torch.arange(W * H * 3, dtype=torch.float32).reshape(H, W, 3)
    slots = s.to_slots(img)
    ok = pix >= 0
    assert slots.shape == (96, 3) and torch.equal(slots[torch.from_numpy(ok)], img.reshape(-1, 3)[pix[ok].astype(np.int64)])
    back = torch.full_like(img, -1.0)
    s.from_slots(slots + 1000.0, back)             # every pixel written exactly once, nothing from the empty slots
    assert torch.equal(back, img + 1000.0)
    assert torch.equal(s.global_pixels(), torch.from_numpy(pix.astype(np.int64)))
    plain = store(None)                            # no permutation: slot q is pixel q
    assert plain.J_shape == plain.slot_J_shape == (H, W, 3) and not plain.needs_slot_copy
    assert plain.c_struct().pix is None and plain.c_struct().pixels == W * H
    band = _lib.Band(1, 2, 2, 0)                   # tiles 1..2 of the image as a band: J arrays are per slot, pix is not passed
    b = store(torch.from_numpy(pix[:64]), band=band)
    assert b.is_band and b.J_shape == b.slot_J_shape == (64, 3) and not b.needs_slot_copy
    assert b.c_struct().pix is None and b.c_struct().pixels == 64
