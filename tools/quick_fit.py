"""Developer probe: us per fit_kernel launch (config-2 shape, 200 chained launches, best of 3) for one or more
builds of the library:  python tools/quick_fit.py [lib.so ...]   (default: the in-tree library).
SUCRE_QUICK_BAND=k restricts the store to the first 1/k of the target's tiles (the per-rank share of a k-GPU run)."""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT)]
from sucre_b200 import _lib, engine  # noqa: E402
from sucre_b200.synth import SyntheticScene  # noqa: E402

V, W, H, iters = 100, 1368, 912, 200
band = int(os.environ.get('SUCRE_QUICK_BAND', '1'))
scene = SyntheticScene(V, W, H, seed=0)
ds = engine.DeviceScene('cuda')
depth, rgb = scene.render_all(device='cuda')
ds.add_views(list(range(V)), [engine.ViewGeom.from_pose(*scene.reference_pose(i)) for i in range(V)], depth, rgb)
store = None
for path in (sys.argv[1:] or [str(_lib.LIB_PATH)]):
    _lib._lib, _lib.LIB_PATH = None, Path(path).resolve()
    if store is None:
        nt = (W * H + 31) // 32
        store = engine.gather(ds, 55, list(range(V)), band=None if band == 1 else _lib.Band.cyclic(nt, 0, band))
    store.workspace = None
    best, params = 1e9, None
    for rep in range(4):
        state = engine.FitState.initial('cuda')
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        engine.fit(store, state, iters)
        e1.record()
        torch.cuda.synchronize()
        if rep:
            best = min(best, e0.elapsed_time(e1) / iters * 1e3)
        params = state.params.cpu().numpy()
    print(f'{Path(path).name:40s} {best:8.2f} us/launch  {store.record_bytes * store.n_obs / best / 1e3:7.1f} GB/s algorithmic '
          f'{store.stream_bytes / best / 1e3:7.1f} GB/s streamed  params {params[:3]}', flush=True)
