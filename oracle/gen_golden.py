"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz by running the UNMODIFIED reference.

Run in the development container (where /root/reference exists):

    python oracle/gen_golden.py [case ...]

The reference (/root/reference/sucre/{sucre,sfm,loader,se3}.py) is imported as-is; the three packages it
imports that are absent from this image (h5py, pycolmap, matplotlib) are replaced by the import stand-ins in
oracle/refshim/.  Scenes come from sucre_b200.synth (seeded); the tiny cases store their inputs too so the
fixtures do not depend on the generator.  What is captured per case:

  * per kept view (HDF5 group order = name-sorted): u1,v1,u2,v2 (int16), d (f32), I (3,n f32) as written by
    sfm.Image.match_images + loader.MatchesFile.prepare_matches, and cP (3,n f32) from load_matches;
  * per Adam iteration: B, beta, gamma after optimizer.step() (hook on torch.optim.Adam.step) and the
    printed cost (parsed from the reference's own log line, sucre.py:150-152);
  * final J (H,W,3) and the saved .pt parameters.

Nothing under tests/ or the product imports this file; the GPU box never sees /root/reference.
"""
from __future__ import annotations

import hashlib
import re
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'oracle' / 'refshim'))
sys.path.insert(0, '/root/reference/sucre')

from sucre_b200.synth import SyntheticScene  # noqa: E402

import h5py  # noqa: E402  (the shim)
import tqdm as _tqdm  # noqa: E402
import sfm as ref_sfm  # noqa: E402  (reference)
import sucre as ref_sucre  # noqa: E402  (reference)

GOLDEN = ROOT / 'tests' / 'golden'


def run_reference(scene: SyntheticScene, target: str, *, closed_form: bool, num_iter: int, min_cover: float,
                  batch_size: int, filter_names=(), image_scale: float = 1.0, light_model: bool = False):
    """Runs reference restore_image on CPU; returns dict of captured arrays."""
    log_lines = []
    history = []
    orig_write = _tqdm.tqdm.write
    orig_step = torch.optim.Adam.step

    def write(msg, *a, **k):
        log_lines.append(msg)

    def step(self, *a, **k):
        out = orig_step(self, *a, **k)
        ps = [p.detach().clone() for g in self.param_groups for p in g['params']]
        # registration order: B, beta, gamma, then (light model) cam2light (6) and sigma (2x2)
        history.append(torch.cat([p.flatten() for p in ps[:5 if light_model else 3]]).numpy())
        return out

    with tempfile.TemporaryDirectory() as tmp:
        tmp = Path(tmp)
        dirs = scene.write(tmp)
        out_dir = tmp / 'out'
        out_dir.mkdir()
        _tqdm.tqdm.write = staticmethod(write)
        ref_sucre.tqdm.write = staticmethod(write)
        torch.optim.Adam.step = step
        try:
            model = ref_sfm.COLMAPModel(model_dir=dirs['model'], image_dir=dirs['images'], depth_dir=dirs['depth'],
                                        image_scale=image_scale)
            image = model[target]
            image_list = [im for im in model.images.values() if im.name not in filter_names]
            h5py._reset()
            ref_sucre.restore_image(image=image, colmap_model=model, output_dir=out_dir, light_model=light_model,
                                    use_closed_form=closed_form, min_cover=min_cover, image_list=image_list,
                                    lr=0.05, num_iter=num_iter, batch_size=batch_size, keep_matches=True,
                                    device='cpu')
        finally:
            _tqdm.tqdm.write = orig_write
            torch.optim.Adam.step = orig_step
        matches_path = (out_dir / target).with_suffix('.h5')
        matches_file = ref_sucre.loader.MatchesFile(matches_path, colmap_model=model)
        data = matches_file.load_matches()
        views = {}
        with h5py.File(matches_path, 'r') as f:
            for (name, group), sample in zip(f.items(), data.data):
                views[name] = {k: group[k][()] for k in ('u1', 'v1', 'u2', 'v2', 'd', 'I')}
                views[name]['cP'] = sample.cP.numpy()
                views[name]['z'] = sample.cP.norm(dim=0).numpy()
        saved = torch.load((out_dir / target).with_suffix('.pt'))
        # geometry exactly as the reference holds/derives it: cam->world pose (sfm.py:219-222), K (sfm.py:204-208),
        # K.inverse() (sfm.py:92) and Pose.inverse() (sfm.py:47)
        poses = {im.name: dict(R=im.pose.R.numpy(), t=im.pose.t.numpy(), K=im.camera.K.numpy(),
                               Kinv=im.camera.K.inverse().numpy(), Ri=im.pose.inverse().R.numpy().copy(),
                               ti=im.pose.inverse().t.numpy(), wh=np.array([im.camera.width, im.camera.height]))
                 for im in model.images.values()}
    cost = [float(re.search(r'cost: ([0-9.e+-]+)', l).group(1)) for l in log_lines if l.startswith('iter:')]
    out = dict(views=views, history=np.stack(history), cost=np.array(cost), J=saved['J'].numpy(),
               B=saved['B'].numpy(), beta=saved['beta'].numpy(), gamma=saved['gamma'].numpy(), poses=poses)
    if light_model:
        out['cam2light'] = saved['cam2light'].numpy()
        out['sigma'] = saved['sigma'].numpy()
    return out


def _hash(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def pack_full(scene: SyntheticScene, target: str, res: dict, extra: dict) -> dict:
    """Everything, for tiny scenes (inputs included)."""
    out = dict(extra)
    out['target'] = target
    out['names'] = np.array([scene.image_name(i) for i in range(scene.n_views)])
    for i in range(scene.n_views):
        depth, rgb = scene.render(i)
        out[f'in_depth_{i}'] = depth.numpy()
        out[f'in_rgb_{i}'] = rgb.numpy()
        q, t = scene.cam_from_world(i)
        out[f'in_q_{i}'] = q
        out[f'in_t_{i}'] = t
        out[f'in_cam_{i}'] = np.array(scene.cams[scene.view_cam[i]], dtype=np.float64)
        for k, a in res['poses'][scene.image_name(i)].items():
            out[f'ref_{k}_{i}'] = a
    out['kept'] = np.array(list(res['views'].keys()))
    for name, v in res['views'].items():
        for k, a in v.items():
            out[f'm_{name}_{k}'] = a
    for k in ('history', 'cost', 'J', 'B', 'beta', 'gamma', 'cam2light', 'sigma'):
        if k in res:
            out[k] = res[k]
    return out


def pack_summary(scene: SyntheticScene, target: str, res: dict, extra: dict, j_stride: int = 16) -> dict:
    """Counts / hashes / trajectories only, for full-size scenes regenerated by sucre_b200.synth."""
    out = dict(extra)
    out['target'] = target
    out['kept'] = np.array(list(res['views'].keys()))
    out['n'] = np.array([len(v['u1']) for v in res['views'].values()])
    out['idx_sha256'] = np.array([_hash(v['u1'], v['v1'], v['u2'], v['v2']) for v in res['views'].values()])
    out['obs_sha256'] = np.array([_hash(v['d'], v['I'], v['z']) for v in res['views'].values()])
    out['sum_u2'] = np.array([int(v['u2'].astype(np.int64).sum()) for v in res['views'].values()])
    out['sum_v2'] = np.array([int(v['v2'].astype(np.int64).sum()) for v in res['views'].values()])
    in_hash = hashlib.sha256()
    for i in range(scene.n_views):
        depth, rgb = scene.render(i)
        in_hash.update(depth.numpy().tobytes())
        in_hash.update(rgb.numpy().tobytes())
    out['inputs_sha256'] = in_hash.hexdigest()
    for i in range(scene.n_views):
        for k, a in res['poses'][scene.image_name(i)].items():
            out[f'ref_{k}_{i}'] = a
    J = res['J']
    out['J_nan_count'] = int(np.isnan(J).any(axis=2).sum())
    out['J_sub'] = J[::j_stride, ::j_stride].copy()
    out['J_stride'] = j_stride
    out['J_nanmean'] = np.nanmean(J.reshape(-1, 3).astype(np.float64), axis=0)
    for k in ('history', 'cost', 'B', 'beta', 'gamma'):
        out[k] = res[k]
    return out


def case_tiny6():
    scene = SyntheticScene(6, 96, 64, seed=0)
    target = 'image0002.png'
    for mode, cf in (('closed', True), ('param', False)):
        res = run_reference(scene, target, closed_form=cf, num_iter=25, min_cover=1e-6, batch_size=2)
        packed = pack_full(scene, target, res, dict(closed_form=cf, num_iter=25, min_cover=1e-6,
                                                    seed=0, width=96, height=64, n_views=6))
        if not cf:  # inputs and matches are those of tiny6_closed.npz; keep the fit outputs only
            packed = {k: v for k, v in packed.items() if not k.startswith(('in_', 'm_', 'ref_'))}
        np.savez_compressed(GOLDEN / f'tiny6_{mode}.npz', **packed)
        print('tiny6', mode, {k: len(v['u1']) for k, v in res['views'].items()}, res['cost'][[0, -1]])


def case_mixed8():
    # two cameras of different size; min_cover high enough to drop weakly overlapping views; one view filtered out
    scene = SyntheticScene(8, 120, 80, seed=3, alt_size=(100, 90), alt_every=3)
    for target in ('image0004.png', 'image0002.png'):
        res = run_reference(scene, target, closed_form=True, num_iter=15, min_cover=0.55, batch_size=3,
                            filter_names=('image0007.png',))
        np.savez_compressed(GOLDEN / f'mixed8_{Path(target).stem}.npz',
                            **pack_full(scene, target, res, dict(closed_form=True, num_iter=15, min_cover=0.55,
                                                                 seed=3, width=120, height=80, n_views=8,
                                                                 alt_w=100, alt_h=90, alt_every=3,
                                                                 filtered=np.array(['image0007.png']))))
        print('mixed8', target, {k: len(v['u1']) for k, v in res['views'].items()}, res['cost'][[0, -1]])


def case_config1():
    # BASELINE.json configs[0]: 20 views 640x480, --image-name image0008.png, closed form, 200 iterations (~5 min)
    scene = SyntheticScene(20, 640, 480, seed=0)
    target = 'image0008.png'
    res = run_reference(scene, target, closed_form=True, num_iter=200, min_cover=1e-6, batch_size=5)
    np.savez_compressed(GOLDEN / 'config1_closed.npz',
                        **pack_summary(scene, target, res, dict(closed_form=True, num_iter=200, min_cover=1e-6,
                                                                seed=0, width=640, height=480, n_views=20)))
    print('config1', int(sum(len(v['u1']) for v in res['views'].values())), res['cost'][[0, -1]],
          res['B'].ravel(), res['beta'].ravel(), res['gamma'].ravel())


def case_config1_param():
    # same scene, default CLI mode (J as an Adam parameter), 60 iterations
    scene = SyntheticScene(20, 640, 480, seed=0)
    target = 'image0008.png'
    res = run_reference(scene, target, closed_form=False, num_iter=60, min_cover=1e-6, batch_size=5)
    np.savez_compressed(GOLDEN / 'config1_param.npz',
                        **pack_summary(scene, target, res, dict(closed_form=False, num_iter=60, min_cover=1e-6,
                                                                seed=0, width=640, height=480, n_views=20)))
    print('config1_param', res['cost'][[0, -1]], res['B'].ravel(), res['beta'].ravel(), res['gamma'].ravel())


def case_scaled8():
    # --image-scale 0.5: colour resampled in float (INTER_AREA), depth nearest, intrinsics scaled (sfm.py:193-199)
    scene = SyntheticScene(8, 192, 128, seed=5)
    target = 'image0004.png'
    for mode, cf in (('closed', True), ('param', False)):
        res = run_reference(scene, target, closed_form=cf, num_iter=15, min_cover=1e-6, batch_size=3, image_scale=0.5)
        packed = pack_full(scene, target, res, dict(closed_form=cf, num_iter=15, min_cover=1e-6, seed=5, width=192,
                                                    height=128, n_views=8, image_scale=0.5))
        if not cf:
            packed = {k: v for k, v in packed.items() if not k.startswith(('in_', 'm_', 'ref_'))}
        np.savez_compressed(GOLDEN / f'scaled8_{mode}.npz', **packed)
        print('scaled8', mode, {k: len(v['u1']) for k, v in res['views'].items()}, res['cost'][[0, -1]])


def case_light6():
    # --light-model (sucre.py:44-46, 54-61): Gaussian light cone + two-leg path, 10 extra Adam parameters
    scene = SyntheticScene(6, 96, 64, seed=0)
    target = 'image0002.png'
    for mode, cf in (('closed', True), ('param', False)):
        res = run_reference(scene, target, closed_form=cf, num_iter=30, min_cover=1e-6, batch_size=2, light_model=True)
        packed = pack_full(scene, target, res, dict(closed_form=cf, num_iter=30, min_cover=1e-6, seed=0, width=96,
                                                    height=64, n_views=6, light_model=True))
        packed = {k: v for k, v in packed.items() if not k.startswith(('in_', 'm_', 'ref_'))}  # inputs = tiny6_closed
        np.savez_compressed(GOLDEN / f'light6_{mode}.npz', **packed)
        print('light6', mode, res['cost'][[0, -1]], res['cam2light'], res['sigma'].ravel())


CASES = dict(scaled8=case_scaled8, light6=case_light6, tiny6=case_tiny6, mixed8=case_mixed8, config1=case_config1, config1_param=case_config1_param)

if __name__ == '__main__':
    GOLDEN.mkdir(parents=True, exist_ok=True)
    torch.manual_seed(0)
    for name in (sys.argv[1:] or ['tiny6', 'mixed8']):
        CASES[name]()
