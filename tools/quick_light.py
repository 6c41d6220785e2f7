"""Developer probe: the light model (--light-model) at config-2 size — time per Adam iteration of light.fit (two kernels
+ the 200-byte read-back + the host-side chain rule and torch.optim.Adam step) and of its two kernels alone."""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT)]
from sucre_b200 import engine, light  # noqa: E402
from sucre_b200.synth import SyntheticScene  # noqa: E402

V, W, H, iters = 100, 1368, 912, 20
scene = SyntheticScene(V, W, H, seed=0)
ds = engine.DeviceScene('cuda')
depth, rgb = scene.render_all(device='cuda')
ds.add_views(list(range(V)), [engine.ViewGeom.from_pose(*scene.reference_pose(i)) for i in range(V)], depth, rgb)
store = engine.gather(ds, 55, list(range(V)), with_points=True)
names = ('B', 'beta', 'gamma', 'cam2light', 'sigma')
params = {'B': torch.full((3, 1), 0.1), 'beta': torch.full((3, 1), 0.1), 'gamma': torch.full((3, 1), 0.1),
          'cam2light': torch.zeros(6), 'sigma': torch.eye(2)}
for p in params.values():
    p.requires_grad_(True)
opt = torch.optim.Adam([params[k] for k in names], lr=0.05)
light.fit(store, params, None, None, 3, 0.05, opt)
torch.cuda.synchronize()
t0 = time.perf_counter()
light.fit(store, params, None, None, iters, 0.05, opt, first_step=4)
torch.cuda.synchronize()
per_iter = (time.perf_counter() - t0) / iters
p24 = light.derive(*(params[k].detach() for k in names)).cuda()
sums = torch.zeros(25, dtype=torch.float64, device='cuda')
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record()
for _ in range(iters):
    J = engine.light_J(store, p24)
e[1].record()
for _ in range(iters):
    engine.light_sums(store, p24, J, sums)
e[2].record()
torch.cuda.synchronize()
tJ, tS = e[0].elapsed_time(e[1]) / iters * 1e3, e[1].elapsed_time(e[2]) / iters * 1e3
b = store.stream_bytes
print(f'light model, config 2: N = {store.n_obs}, {store.record_bytes} B records, {b / 1e6:.0f} MB per sweep (fill {store.fill:.3f})')
print(f'  light.fit: {per_iter * 1e6:.0f} us per iteration wall (two sweeps + 200-byte read-back + host chain rule / Adam)')
print(f'  light_J_kernel    {tJ:7.1f} us  {b / tJ / 1e3:6.0f} GB/s streamed')
print(f'  light_sums_kernel {tS:7.1f} us  {b / tS / 1e3:6.0f} GB/s streamed   (+ its 25-column reduction kernel)')
