"""Pins oracle/torch_port.py (the CPU baseline timed by bench.py) bit-for-bit to the unmodified reference's outputs."""
import numpy as np
import pytest
import torch

from oracle import torch_port as tp


def _views(g):
    out = {}
    for i, name in enumerate(g['names'].tolist()):
        a = g.geom_arrays(i)
        depth, rgb = g.inputs(i)
        t = lambda x: torch.tensor(np.asarray(x), dtype=torch.float32)  # noqa: E731
        out[name] = tp.make_view(t(a['K']), t(a['R']), t(a['t']).reshape(3, 1), a['wh'][0], a['wh'][1],
                                 torch.from_numpy(depth.astype(np.int32)), torch.from_numpy(rgb.copy()))
    return out


@pytest.mark.parametrize('case', ['tiny6_closed', 'mixed8_image0004'])
def test_port_gather_and_fit_match_reference(golden, case):
    g = golden(case)
    views = _views(g)
    names = g['names'].tolist()
    sources = [(names[i], views[names[i]]) for i in g.pairing_list()]
    kept = tp.gather(views[str(g['target'])], sources, min_cover=float(g['min_cover']))
    assert [k for k, _ in kept] == g['kept'].tolist()
    for name, obs in kept:
        ref = g.matches(name)
        for mine, key in (('u', 'u1'), ('v', 'v1'), ('u2', 'u2'), ('v2', 'v2'), ('d', 'd'), ('cP', 'cP'), ('I', 'I')):
            assert np.array_equal(obs[mine].numpy(), ref[key]), (name, key)
    W, H = g.geom_arrays(g.view_index(str(g['target'])))['wh']
    model = tp.FormationModel(int(H), int(W), closed_form=True)
    hist, cost = tp.run_adam(model, [o for _, o in kept], int(g['num_iter']), batch_size=2 if case.startswith('tiny') else 3)
    assert np.array_equal(hist.numpy(), g['history'])  # same ops, same order, same machine => same bits
    assert np.allclose(cost.numpy(), g['cost'], rtol=1e-4)
    assert np.array_equal(np.isnan(model.J.numpy()), np.isnan(g['J']))
    assert np.nanmax(np.abs(model.J.numpy() - g['J'])) == 0
