"""Developer probe: the margins of tests/test_fullsize_gpu.py::test_fit_properties_at_full_size (kernel sums against an
independent float64 evaluation, stationarity of the closed-form J) without the oracle legs — a few seconds on the GPU."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / 'tests')]
import helpers  # noqa: E402
from sucre_b200 import engine  # noqa: E402
from sucre_b200.synth import SyntheticScene  # noqa: E402

V, W, H, TARGET = 100, 1368, 912, 55
scene = SyntheticScene(V, W, H, seed=0)
ds = engine.DeviceScene('cuda')
depth, rgb = scene.render_all(device='cuda')
ds.add_views(list(range(V)), [engine.ViewGeom.from_pose(*helpers.reference_pose(scene, i)) for i in range(V)], depth, rgb)
store = engine.gather(ds, TARGET, list(range(V)))
state = engine.FitState.initial(ds.device)
sums = torch.zeros(10, dtype=torch.float64, device=ds.device)
first = torch.zeros_like(sums)
engine.fit_sums(store, state, first)
engine.fit_sums(store, state, sums)
J = engine.closed_form_J(store, state.params).reshape(-1, 3)
cell, pixel, _ = store.record_index()
rec = store.cells[cell].double()
z, I = rec[:, :1], rec[:, 1:]
B, beta, gamma = (state.params[i:i + 3].double() for i in (0, 3, 6))
a, e = torch.exp(-beta * z), torch.exp(-gamma * z)
Jp = J.double()[pixel]
r = I - (Jp * a + B * (1 - e))
stat = torch.zeros((W * H, 3), dtype=torch.float64, device=ds.device).index_add_(0, pixel, r * a)
norm = torch.zeros((W * H, 3), dtype=torch.float64, device=ds.device).index_add_(0, pixel, (I * a).abs())
ref = torch.cat([(r * (1 - e)).sum(0), (r * Jp * z * a).sum(0), (r * B * z * e).sum(0), (r * r).sum().reshape(1)])
scale = torch.cat([(r * (1 - e)).abs().sum(0), (r * Jp * z * a).abs().sum(0), (r * B * z * e).abs().sum(0),
                   (r * r).sum().reshape(1)])
print(f'stationarity {float((stat.abs() / norm.clamp_min(1e-30)).max()):.3e} (bound 2e-6)')
print(f'sums, J_ref = 0 {float(((first - ref).abs() / scale).max()):.3e} (bound 2e-5)')
print(f'sums, J_ref = previous J {float(((sums - ref).abs() / scale).max()):.3e} (bound 1e-5)')
