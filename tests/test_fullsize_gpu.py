"""Parity at BASELINE.json's full size (configs[1]: 100 views, 1368x912, one target): every one of the 125 M
pixel-views against the oracle, plus size-independent properties of the fit on the full observation store."""
import numpy as np
import pytest
import torch

from oracle import oracle
from sucre_b200 import api, engine
from sucre_b200.synth import SyntheticScene

import helpers

pytestmark = pytest.mark.gpu

V, W, H, TARGET = 100, 1368, 912, 55


@pytest.fixture(scope='module')
def full():
    scene = SyntheticScene(V, W, H, seed=0)
    ds, host = helpers.build_device_scene(scene, range(V), render_device='cuda')
    store = engine.gather(ds, TARGET, list(range(V)), keep_src=True)
    return ds, host, store


def test_every_pixel_view_matches_the_oracle(full):
    """Mask and integer source pixel of all 100 views x 1 247 616 target pixels: bit-exact, 0 exceptions."""
    ds, host, store = full
    cell, pixel, view = store.record_index()
    src = store.cell_src[cell]
    n_in_total = 0
    for s in range(V):
        idx, n, n_in = oracle.match_pair(host[TARGET][0], host[TARGET][2], host[s][0], host[s][2])
        n_in_total += n_in
        assert n == store.view_count[s], (s, n, store.view_count[s])
        if not store.view_kept[s]:
            assert not n / (W * H) > 1e-6
            continue
        sel = view == s
        mine = torch.full((W * H,), -1, dtype=torch.int32, device=ds.device)
        mine[pixel[sel]] = src[sel]
        assert np.array_equal(mine.cpu().numpy(), idx.reshape(-1)), f'view {s}'
    assert store.n_obs == int(store.view_count[store.view_kept].sum())
    # the match kernel's own statistics (bench.py's roofline_gather): in-bounds forward projections counted on the
    # device equal the oracle's, and the frustum pre-test skipped (tile, view) pairs without changing any mask
    assert store.stats['n_inbounds'] == n_in_total, (store.stats, n_in_total)
    assert 0 < store.stats['tile_views_culled'] < store.stats['tile_views']
    print(f'config 2: {store.n_obs} observations, {int(store.view_kept.sum())}/{V} views kept, '
          f'{n_in_total} in-bounds forward projections')


def test_payload_matches_the_oracle_on_sampled_views(full):
    ds, host, store = full
    got = store.to_reference_layout()
    for s in (TARGET, 44, 66, 3):
        if not store.view_kept[s]:
            continue
        idx, _, _ = oracle.match_pair(host[TARGET][0], host[TARGET][2], host[s][0], host[s][2])
        ref = oracle.sample_pair(idx, host[s][0], host[s][1], host[s][2])
        for f in ('u1', 'v1', 'u2', 'v2'):
            assert np.array_equal(got[s][f], ref[f]), (s, f)
        assert np.array_equal(got[s]['z'].view(np.uint32), ref['z'].view(np.uint32))
        assert np.array_equal(got[s]['I'].view(np.uint32), ref['I'].view(np.uint32))


def test_store_structure_invariants(full):
    """Reference invariants (loader.py:89-101) and the mutual-match property at full size."""
    ds, host, store = full
    rec = store.records()
    assert not torch.isnan(rec).any() and (rec[:, 0] > 0).all() and (rec[:, 1:] >= 0).all() and (rec[:, 1:] <= 1).all()
    cell, pixel, view = store.record_index()
    assert torch.unique(cell).numel() == store.n_obs                       # every record slot used exactly once
    assert store.n_rows * 32 >= store.n_obs and 0.97 < store.fill <= 1.0   # slots dealt by count: few sentinels (0.90 without)
    pix = store.pix.cpu().numpy()
    assert np.array_equal(np.sort(pix[pix >= 0]), np.arange(W * H))         # every pixel sits in exactly one slot
    key = view * (W * H) + pixel
    assert torch.unique(key).numel() == store.n_obs                        # a pixel is matched at most once per view
    srckey = view * (1 << 32) + store.cell_src[cell].to(torch.int64) % (1 << 32)
    assert torch.unique(srckey).numel() == store.n_obs                     # ... and a source pixel at most once
    sel = view == TARGET                                                   # self-match = identity on valid pixels
    u2 = store.cell_src[cell[sel]] & 0xffff
    v2 = (store.cell_src[cell[sel]] >> 16) & 0xffff
    assert torch.equal(v2.to(torch.int64) * W + u2, pixel[sel])
    assert int(sel.sum()) == int((host[TARGET][0] > 0).sum())


def test_fit_properties_at_full_size(full):
    """Closed-form J is a stationary point of the per-pixel problem (sum r a = 0), the sums reported by the kernel
    equal an independent float64 evaluation over the exported records, and a full 200-iteration run agrees with
    the oracle's parameters to 1e-4."""
    ds, host, store = full
    state = engine.FitState.initial(ds.device)
    sums = torch.zeros(10, dtype=torch.float64, device=ds.device)
    first = torch.zeros_like(sums)
    engine.fit_sums(store, state, first)   # reference point J_ref = 0: the statistics still cancel (~1e-5)
    engine.fit_sums(store, state, sums)    # reference point = J of the previous evaluation: residual-scale products
    J = engine.closed_form_J(store, state.params).reshape(-1, 3)
    _, pixel, _ = store.record_index()
    rec = store.records().double()
    z, I = rec[:, :1], rec[:, 1:]
    B, beta, gamma = (state.params[i:i + 3].double() for i in (0, 3, 6))
    a, e = torch.exp(-beta * z), torch.exp(-gamma * z)
    Jp = J.double()[pixel]
    r = I - (Jp * a + B * (1 - e))
    stat = torch.zeros((W * H, 3), dtype=torch.float64, device=ds.device).index_add_(0, pixel, r * a)
    norm = torch.zeros((W * H, 3), dtype=torch.float64, device=ds.device).index_add_(0, pixel, (I * a).abs())
    assert float((stat.abs() / norm.clamp_min(1e-30)).max()) < 2e-6        # J solves its normal equation
    ref = torch.cat([(r * (1 - e)).sum(0), (r * Jp * z * a).sum(0), (r * B * z * e).sum(0), (r * r).sum().reshape(1)])
    scale = torch.cat([(r * (1 - e)).abs().sum(0), (r * Jp * z * a).abs().sum(0), (r * B * z * e).abs().sum(0),
                       (r * r).sum().reshape(1)])
    # fp32 statistics + ex2.approx (<= 2^-22 relative) against exact float64 exponentials
    assert float(((first - ref).abs() / scale).max()) < 2e-5
    assert float(((sums - ref).abs() / scale).max()) < 1e-5
    # full run vs oracle (CPU, OpenMP): same observations, 200 iterations
    res = api.restore_resident(ds, TARGET, list(range(V)), num_iter=200)
    kept, _ = helpers.oracle_gather(host, TARGET, [s for s in range(V) if store.view_kept[s]])
    oref = oracle.fit([o for _, o in kept], W, H, closed_form=True, num_iter=200)
    p = res.params.cpu().numpy()
    assert np.max(np.abs(p - oref['params']) / np.abs(oref['params'])) < 1e-4
    Jg = res.J.cpu().numpy()
    assert np.array_equal(np.isnan(Jg), np.isnan(oref['J'])) and np.nanmax(np.abs(Jg - oref['J'])) < 1e-3


def test_param_mode_at_full_size_vs_oracle(full):
    """The default CLI mode (J is an Adam parameter, sucre.py:47-50) at BASELINE configs[1] size: 40 iterations of the
    fused kernel (per-pixel Adam step of J inside the sweep) against the oracle on the same observations."""
    ds, host, store = full
    kept, _ = helpers.oracle_gather(host, TARGET, [s for s in range(V) if store.view_kept[s]])
    J0 = oracle.initial_J(host[TARGET][1], host[TARGET][0])
    oref = oracle.fit([o for _, o in kept], W, H, closed_form=False, num_iter=40, J0=J0)
    state = engine.FitState.initial(ds.device, J0=torch.from_numpy(J0))
    hist = engine.fit(store, state, 40).cpu().numpy()
    p = state.params.cpu().numpy()
    assert np.max(np.abs(p - oref['params']) / np.abs(oref['params'])) < 1e-4
    assert np.max(np.abs(hist[:, :9] - oref['history']) / np.maximum(np.abs(oref['history']), 0.05)) < 1e-4
    assert np.max(np.abs(hist[:, 9] - oref['cost']) / oref['cost']) < 1e-5
    Jg = state.J.cpu().numpy()
    assert np.array_equal(np.isnan(Jg), np.isnan(oref['J'])) and np.nanmax(np.abs(Jg - oref['J'])) < 1e-3


def test_light_model_sums_at_full_size_against_float64():
    """--light-model at BASELINE configs[1] size: the store with camera-frame points (u8 colour, 16-byte records), the
    closed-form J with the light terms and the 25 reduced sums of one residual pass against an independent float64
    evaluation of sucre.py:54-61, 66-82 over the exported records."""
    from sucre_b200 import light
    scene = SyntheticScene(V, W, H, seed=0)
    ds, host = helpers.build_device_scene(scene, range(V), render_device='cuda')
    store = engine.gather(ds, TARGET, list(range(V)), with_points=True)
    assert store.has_points and store.record_bytes == 16 and store.n_obs > 3e7
    dev = ds.device
    B, beta, gamma = (torch.tensor(x, dtype=torch.float32).view(3, 1) for x in ((0.07, 0.2, 0.3), (0.4, 0.15, 0.1), (0.3, 0.2, 0.12)))
    cam2light = torch.tensor([0.02, -0.03, 0.01, 0.05, -0.04, 0.02], dtype=torch.float32)
    sigma = torch.tensor([[0.9, 0.1], [-0.05, 1.1]], dtype=torch.float32)
    p24 = light.derive(B, beta, gamma, cam2light, sigma).to(dev)
    J = engine.light_J(store, p24)
    sums = torch.zeros(25, dtype=torch.float64, device=dev)
    engine.light_sums(store, p24, J, sums)
    # float64 evaluation
    slot, pixel, _ = store.record_index()
    flat = store.cells.reshape(-1, 4)
    cP = flat[slot][:, :3].double()
    I = torch.from_numpy(store.colours(slot)).to(dev).double()
    q = p24.double()
    Bq, bq, gq, R, t, S = q[0:3], q[3:6], q[6:9], q[9:18].view(3, 3), q[18:21], q[21:24]
    lP = cP @ R.T + t
    nl = lP.norm(dim=1)
    x, y = lP[:, 0] / lP[:, 2], lP[:, 1] / lP[:, 2]
    l = torch.exp(-0.5 * (S[0] * x * x + 2 * S[1] * x * y + S[2] * y * y))[:, None]
    z = (cP.norm(dim=1) + nl)[:, None]
    a, e = torch.exp(-bq * z), torch.exp(-gq * z)
    absorb, back = l * a, l * Bq * (1 - e)
    num = torch.zeros((W * H, 3), dtype=torch.float64, device=dev).index_add_(0, pixel, (I - back) * absorb)
    den = torch.zeros((W * H, 3), dtype=torch.float64, device=dev).index_add_(0, pixel, absorb * absorb)
    J64 = num / den
    Jf = J.reshape(-1, 3)
    assert torch.equal(torch.isnan(J64), torch.isnan(Jf)) and float((Jf.double() - J64).nan_to_num(0.0).abs().max()) < 1e-4
    Jp = Jf.double()[pixel]                                  # the residual pass reads the fp32 J the kernel wrote
    M = Jp * a + Bq * (1 - e)
    r = I - l * M
    g_l = (r * M).sum(1)
    g_z = (r * l * (-bq * Jp * a + gq * Bq * e)).sum(1)
    dlx, dly = -l[:, 0] * (S[0] * x + S[1] * y), -l[:, 0] * (S[1] * x + S[2] * y)
    iz = 1.0 / lP[:, 2]
    gP = torch.stack([g_l * dlx * iz + g_z * lP[:, 0] / nl, g_l * dly * iz + g_z * lP[:, 1] / nl,
                      -g_l * (dlx * x + dly * y) * iz + g_z * lP[:, 2] / nl], dim=1)
    gll = g_l * l[:, 0]
    terms = [r * l * (1 - e), r * l * Jp * z * a, r * l * Bq * z * e, (r * r).sum(1, keepdim=True),
             torch.stack([gll * x * x, gll * x * y, gll * y * y], dim=1),
             (gP[:, :, None] * cP[:, None, :]).reshape(-1, 9), gP]
    ref = torch.cat([t_.sum(0) for t_ in terms])
    scale = torch.cat([t_.abs().sum(0) for t_ in terms])
    assert float(((sums - ref).abs() / scale).max()) < 2e-5
