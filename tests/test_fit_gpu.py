"""Parity of the CUDA fit (through the C ABI) with the oracle and the reference's trajectories.
Tolerances are BASELINE.json's: J within 1e-3 max-abs, B / beta / gamma within 1e-4 relative."""
import numpy as np
import pytest
import torch

from oracle import oracle
from sucre_b200 import engine
from sucre_b200.synth import SyntheticScene

import helpers

pytestmark = pytest.mark.gpu

J_ATOL = 1e-3
P_RTOL = 1e-4


def _rel(a, b, floor=1e-12):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


@pytest.mark.parametrize('case', ['tiny6_closed', 'mixed8_image0004', 'mixed8_image0002'])
def test_closed_form_fit_vs_reference_golden(golden, case):
    g = golden(case)
    ds, _ = helpers.golden_device_scene(g)
    names = g['names'].tolist()
    order = sorted((names[i] for i in g.pairing_list()))
    store = engine.gather(ds, str(g['target']), order, min_cover=float(g['min_cover']))
    state = engine.FitState.initial(ds.device)
    hist = engine.fit(store, state, int(g['num_iter'])).cpu().numpy()
    J = engine.closed_form_J(store, state.params).cpu().numpy()
    ref_p = np.concatenate([g['B'].ravel(), g['beta'].ravel(), g['gamma'].ravel()])
    assert _rel(state.params.cpu().numpy(), ref_p) < P_RTOL
    assert _rel(hist[:, :9], g['history'], floor=0.05) < P_RTOL      # whole trajectory; params cross zero mid-run
    assert _rel(hist[:, 9], g['cost']) < 2e-4                        # reference prints 5 significant digits
    assert np.array_equal(np.isnan(J), np.isnan(g['J']))             # identical NaN set
    assert np.nanmax(np.abs(J - g['J'])) < J_ATOL


def test_closed_form_fit_vs_oracle_config1():
    """configs[0] shape, 40 iterations, against the oracle on the same observations."""
    scene = SyntheticScene(20, 640, 480, seed=0)
    ds, host = helpers.build_device_scene(scene, range(20), render_device='cuda')
    store = engine.gather(ds, 8, list(range(20)))
    kept, _ = helpers.oracle_gather(host, 8, list(range(20)))
    ref = oracle.fit([o for _, o in kept], 640, 480, closed_form=True, num_iter=40)
    state = engine.FitState.initial(ds.device)
    hist = engine.fit(store, state, 40).cpu().numpy()
    J = engine.closed_form_J(store, state.params).cpu().numpy()
    assert _rel(state.params.cpu().numpy(), ref['params']) < P_RTOL
    assert _rel(hist[:, :9], ref['history'], floor=0.05) < P_RTOL
    assert _rel(hist[:, 9], ref['cost']) < 1e-5
    assert np.array_equal(np.isnan(J), np.isnan(ref['J']))
    assert np.nanmax(np.abs(J - ref['J'])) < J_ATOL


def test_param_mode_fit_vs_reference_golden(golden):
    """Default CLI mode: J is an Adam parameter (sucre.py:47-50), fused per-pixel Adam in the fit kernel."""
    g, gp = golden('tiny6_closed'), golden('tiny6_param')
    ds, _ = helpers.golden_device_scene(g)
    store = engine.gather(ds, str(g['target']), sorted(g['names'].tolist()))
    depth, rgb = g.inputs(g.view_index(str(g['target'])))
    state = engine.FitState.initial(ds.device, J0=torch.from_numpy(oracle.initial_J(rgb, depth)))
    hist = engine.fit(store, state, int(gp['num_iter'])).cpu().numpy()
    ref_p = np.concatenate([gp['B'].ravel(), gp['beta'].ravel(), gp['gamma'].ravel()])
    assert _rel(state.params.cpu().numpy(), ref_p) < P_RTOL
    assert _rel(hist[:, :9], gp['history'], floor=0.05) < P_RTOL
    assert _rel(hist[:, 9], gp['cost']) < 2e-4
    J = state.J.cpu().numpy()
    assert np.array_equal(np.isnan(J), np.isnan(gp['J']))
    assert np.nanmax(np.abs(J - gp['J'])) < J_ATOL


def test_param_mode_fit_vs_oracle_unobserved_pixels():
    """Target outside its own pairing list: valid pixels without any observation keep their initial J, invalid
    ones stay NaN (zero gradient, sucre.py:49)."""
    scene = SyntheticScene(6, 128, 96, seed=7)
    ds, host = helpers.build_device_scene(scene, range(6))
    src = [0, 1, 2, 4, 5]
    store = engine.gather(ds, 3, src)
    kept, _ = helpers.oracle_gather(host, 3, src)
    J0 = oracle.initial_J(host[3][1], host[3][0])
    ref = oracle.fit([o for _, o in kept], 128, 96, closed_form=False, num_iter=30, J0=J0)
    state = engine.FitState.initial(ds.device, J0=torch.from_numpy(J0))
    hist = engine.fit(store, state, 30).cpu().numpy()
    J = state.J.cpu().numpy()
    assert _rel(state.params.cpu().numpy(), ref['params']) < P_RTOL
    assert _rel(hist[:, 9], ref['cost']) < 1e-5
    assert np.array_equal(np.isnan(J), np.isnan(ref['J'])) and np.isnan(J).any()
    assert np.nanmax(np.abs(J - ref['J'])) < J_ATOL
    untouched = ~np.isnan(J0).any(axis=2)
    seen = np.zeros((96, 128), bool)
    for _, o in kept:
        seen[o['v1'], o['u1']] = True
    untouched &= ~seen
    assert untouched.any() and np.array_equal(J[untouched], J0[untouched])


def test_fit_building_blocks_match_fused_loop(golden):
    """sums + adam_step (the multi-GPU building blocks) reproduce the fused single-GPU loop bit for bit,
    and splitting a run in two (state carried over) changes nothing."""
    g = golden('tiny6_closed')
    ds, _ = helpers.golden_device_scene(g)
    store = engine.gather(ds, str(g['target']), sorted(g['names'].tolist()))
    a = engine.FitState.initial(ds.device)
    ha = engine.fit(store, a, 12)
    b = engine.FitState.initial(ds.device)
    sums = torch.zeros(10, dtype=torch.float64, device=ds.device)
    rows = torch.zeros((12, 10), dtype=torch.float32, device=ds.device)
    for it in range(12):
        engine.fit_sums(store, b, sums)
        engine.adam_step(b, sums, store.n_obs, 0.05, rows[it])
    assert torch.equal(a.params, b.params) and torch.equal(ha, rows) and a.step == b.step == 12
    c = engine.FitState.initial(ds.device)
    hc = torch.cat([engine.fit(store, c, 5), engine.fit(store, c, 7)])
    assert torch.equal(a.params, c.params) and torch.equal(ha, hc)


def test_fit_is_deterministic(golden):
    g = golden('mixed8_image0004')
    ds, _ = helpers.golden_device_scene(g)
    names = g['names'].tolist()
    store = engine.gather(ds, str(g['target']), sorted(names[i] for i in g.pairing_list()), min_cover=float(g['min_cover']))
    runs = []
    for _ in range(2):
        s = engine.FitState.initial(ds.device)
        runs.append((engine.fit(store, s, 10).clone(), s.params.clone()))
    assert torch.equal(runs[0][0], runs[1][0]) and torch.equal(runs[0][1], runs[1][1])


def test_recovers_ground_truth_parameters():
    """Size-independent property: on noise-free synthetic data the fit moves toward the generating model —
    the cost falls by > 10x and closed-form J stays within the texture's range on observed pixels."""
    scene = SyntheticScene(9, 160, 120, seed=2)
    ds, _ = helpers.build_device_scene(scene, range(9))
    store = engine.gather(ds, 4, list(range(9)))
    state = engine.FitState.initial(ds.device)
    hist = engine.fit(store, state, 200).cpu().numpy()
    assert hist[-1, 9] < hist[0, 9] / 5
    J = engine.closed_form_J(store, state.params).cpu().numpy()
    assert np.nanmin(J) > -0.5 and np.nanmax(J) < 1.5


@pytest.mark.parametrize('shape', [(40, 96, 64, 20), (6, 257, 131, 2), (12, 64, 48, 5), (3, 33, 12, 1)])
def test_every_partition_shape_against_float64(shape):
    """The fit splits rows evenly over 148 x 16 warps, cutting tiles between neighbouring warps.  Shapes that stress
    it: tiles far longer than a warp's share (40 views on a small image), ragged last tiles, stores smaller than the
    grid, and a band of empty tiles (target depth zeroed).  Checked: the kernel's sums against an independent float64
    evaluation over the exported records, J against the float64 closed form, and a short J-parameter run against the
    oracle."""
    V, W, H, target = shape
    scene = SyntheticScene(V, W, H, seed=13)
    ds, host = helpers.build_device_scene(scene, range(V))
    if H > 16:  # rows 4..11 of the target become invalid: empty tiles in the middle of the store
        ds.depth[target].view(torch.int16)[4:12] = 0
        host[target][0][4:12] = 0
    keys = list(range(V))
    store = engine.gather(ds, target, keys)
    assert store.n_obs > 0 and store.n_rows * 32 >= store.n_obs
    state = engine.FitState.initial(ds.device)
    sums = torch.zeros(10, dtype=torch.float64, device=ds.device)
    engine.fit_sums(store, state, sums)
    engine.fit_sums(store, state, sums)    # second evaluation: reference point = J of the first
    J = engine.closed_form_J(store, state.params).reshape(-1, 3)
    _, pixel, _ = store.record_index()
    rec = store.records().double()
    z, I = rec[:, :1], rec[:, 1:]
    B, beta, gamma = (state.params[i:i + 3].double() for i in (0, 3, 6))
    a, e = torch.exp(-beta * z), torch.exp(-gamma * z)
    num = torch.zeros((W * H, 3), dtype=torch.float64, device=ds.device).index_add_(0, pixel, (I - B * (1 - e)) * a)
    den = torch.zeros((W * H, 3), dtype=torch.float64, device=ds.device).index_add_(0, pixel, a * a)
    J64 = num / den                                                       # 0/0 = NaN where unobserved (sucre.py:77)
    assert torch.equal(torch.isnan(J64), torch.isnan(J))
    assert float((J.double() - J64).nan_to_num(0.0).abs().max()) < 1e-5
    Jp = J64[pixel]
    r = I - (Jp * a + B * (1 - e))
    ref = torch.cat([(r * (1 - e)).sum(0), (r * Jp * z * a).sum(0), (r * B * z * e).sum(0), (r * r).sum().reshape(1)])
    scale = torch.cat([(r * (1 - e)).abs().sum(0), (r * Jp * z * a).abs().sum(0), (r * B * z * e).abs().sum(0),
                       (r * r).sum().reshape(1)])
    assert float(((sums - ref).abs() / scale).max()) < 1e-5
    # J-parameter mode on the same store: 5 iterations against the oracle
    kept, _ = helpers.oracle_gather(host, target, keys)
    J0 = oracle.initial_J(host[target][1], host[target][0])
    oref = oracle.fit([o for _, o in kept], W, H, closed_form=False, num_iter=5, J0=J0)
    sp = engine.FitState.initial(ds.device, J0=torch.from_numpy(J0))
    hist = engine.fit(store, sp, 5).cpu().numpy()
    assert _rel(sp.params.cpu().numpy(), oref['params']) < P_RTOL and _rel(hist[:, 9], oref['cost']) < 1e-5
    Jg = sp.J.cpu().numpy()
    assert np.array_equal(np.isnan(Jg), np.isnan(oref['J'])) and np.nanmax(np.abs(Jg - oref['J'])) < J_ATOL


def test_empty_store_raises():
    scene = SyntheticScene(2, 64, 48, seed=1)
    ds, _ = helpers.build_device_scene(scene, range(2))
    store = engine.gather(ds, 0, [0, 1], min_cover=1.0)
    with pytest.raises(engine._lib.SucreError):
        engine.fit(store, engine.FitState.initial(ds.device), 3)
    assert torch.isnan(engine.closed_form_J(store, engine.FitState.initial(ds.device).params)).all()
