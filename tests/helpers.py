"""Shared glue of the parity tests: the same inputs are fed to the CUDA path (sucre_b200) and to the oracle."""
from __future__ import annotations

import numpy as np
import torch

from oracle import oracle
from sucre_b200.engine import DeviceScene, ViewGeom
from sucre_b200.synth import SyntheticScene


def geoms_from_golden(g):
    """ViewGeom + OracleView per view index, taken verbatim from the fixture (no host LAPACK involved)."""
    vg, og = {}, {}
    for i in range(g.n_views):
        a = g.geom_arrays(i)
        t = lambda x: torch.tensor(np.asarray(x), dtype=torch.float32)  # noqa: E731
        vg[i] = ViewGeom(K=t(a['K']), Kinv=t(a['Kinv']), R=t(a['R']), t=t(a['t']).reshape(3, 1), Ri=t(a['Ri']),
                         ti=t(a['ti']).reshape(3, 1), width=int(a['wh'][0]), height=int(a['wh'][1]))
        og[i] = oracle.view_geom(a['K'], a['R'], a['t'], a['wh'][0], a['wh'][1], Kinv=a['Kinv'], Ri=a['Ri'], ti=a['ti'])
    return vg, og


def golden_device_scene(g, device='cuda'):
    """DeviceScene keyed by image name holding the fixture's inputs, plus the oracle-side sources."""
    vg, og = geoms_from_golden(g)
    names = g['names'].tolist()
    ds = DeviceScene(device)
    host = {}
    for i, name in enumerate(names):
        depth, rgb = g.inputs(i)
        ds.add_view(name, vg[i], torch.from_numpy(depth.copy()), torch.from_numpy(rgb.copy()))
        host[name] = (depth, rgb, og[i])
    return ds, host


def reference_pose(scene: SyntheticScene, i: int):
    """K, R, t (cam->world), width, height as the reference's COLMAPModel derives them (SyntheticScene.reference_pose)."""
    return scene.reference_pose(i)


def build_device_scene(scene: SyntheticScene, views, device='cuda', render_device=None):
    """Renders `views` and returns (DeviceScene keyed by view index, {i: (depth np, rgb np, OracleView)})."""
    ds = DeviceScene(device)
    host = {}
    for i in views:
        K, R, t, W, H = reference_pose(scene, i)
        geom = ViewGeom.from_pose(K, R, t, W, H)
        depth, rgb = scene.render(i, device=render_device or 'cpu')
        ds.add_view(i, geom, depth, rgb)
        host[i] = (depth.cpu().numpy(), rgb.cpu().numpy(),
                   oracle.view_geom(geom.K, geom.R, geom.t, W, H, Kinv=geom.Kinv, Ri=geom.Ri, ti=geom.ti))
    return ds, host


def oracle_gather(host, target_key, source_keys, min_cover=1e-6, sample=True):
    sources = [(k, *host[k]) for k in source_keys]
    return oracle.gather(host[target_key][0], host[target_key][2], sources, min_cover=min_cover, sample=sample)


def compare_store_with_oracle(store, kept_oracle) -> dict:
    """store: ObservationStore (CUDA); kept_oracle: [(key, obs dict)] from oracle.gather.
    Returns mismatch counts; kept-view lists and per-view sizes must agree outright."""
    got = store.to_reference_layout()
    assert list(got.keys()) == [k for k, _ in kept_oracle], (list(got.keys()), [k for k, _ in kept_oracle])
    bad = dict(idx=0, z=0, I=0, n=0)
    for key, ref in kept_oracle:
        mine = got[key]
        assert mine['u1'].shape == ref['u1'].shape, (key, mine['u1'].shape, ref['u1'].shape)
        bad['n'] += len(ref['u1'])
        for f in ('u1', 'v1', 'u2', 'v2'):
            bad['idx'] += int((mine[f] != ref[f]).sum())
        bad['z'] += int((mine['z'].view(np.uint32) != ref['z'].view(np.uint32)).sum())
        bad['I'] += int((mine['I'].view(np.uint32) != ref['I'].view(np.uint32)).sum())
    return bad
