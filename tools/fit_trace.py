"""Developer probe: where the time of one Adam iteration goes inside the resident fit kernel.

    tools/fit_variants.sh "trace -DSUCRE_FIT_TRACE"
    gpurun -- 'SUCRE_QUICK_BAND=8 python tools/fit_trace.py variants/trace.so'

The trace build stamps %globaltimer per CTA at: 0 iteration top, 1 all warps of the CTA done with their sweep,
2 row published, 3 warp 0 has all rows, 4 whole CTA has the sums, 5 Adam step done; and per warp at the end of its sweep.
Prints, over iterations 8 .. 63, the mean length of every phase and the spread over the CTAs / warps."""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT)]
from sucre_b200 import _lib, engine  # noqa: E402
from sucre_b200.synth import SyntheticScene  # noqa: E402

V, W, H, iters = 100, 1368, 912, 64
band = int(os.environ.get('SUCRE_QUICK_BAND', '1'))
_lib._lib, _lib.LIB_PATH = None, Path(sys.argv[1]).resolve()
scene = SyntheticScene(V, W, H, seed=0)
ds = engine.DeviceScene('cuda')
depth, rgb = scene.render_all(device='cuda')
ds.add_views(list(range(V)), [engine.ViewGeom.from_pose(*scene.reference_pose(i)) for i in range(V)], depth, rgb)
nt = (W * H + 31) // 32
store = engine.gather(ds, 55, list(range(V)), band=None if band == 1 else _lib.Band.cyclic(nt, 0, band))
for _ in range(3):
    state = engine.FitState.initial('cuda')
    engine.fit(store, state, iters)
torch.cuda.synchronize()
lib = C.CDLL(str(_lib.LIB_PATH))
n_warps = 16
cta = np.zeros((64, 160, 8), np.uint64)
wrp = np.zeros((64, 160, n_warps), np.uint64)
assert lib.sucre_debug_fit_trace(C.c_void_p(cta.ctypes.data), C.c_void_p(wrp.ctypes.data)) == 0
n = torch.cuda.get_device_properties(0).multi_processor_count
cta, wrp = cta[8:, :n].astype(np.int64), wrp[8:, :n].astype(np.int64)
t0 = cta[:, :, 0].min(axis=1, keepdims=True)          # first CTA to start the iteration
names = ['iteration top', 'sweeps of the CTA done (+ warp trees)', 'row published', 'warp 0 has all rows', 'CTA has the sums',
         'Adam done']
print(f'band 1/{band}: {store.n_rows} rows, {store.n_tiles} tiles; ns relative to the first CTA entering the iteration; '
      f'iteration = {np.diff(cta[:, 0, 0]).mean():.0f} ns; timer step = {np.min(np.diff(np.unique(cta))[np.diff(np.unique(cta)) > 0])} ns')
for k, name in enumerate(names):
    rel = cta[:, :, k] - t0
    print(f'  {k} {name:40s} mean {rel.mean():8.0f}   earliest CTA {rel.min(axis=1).mean():8.0f}   latest CTA {rel.max(axis=1).mean():8.0f}')
w = wrp - t0[:, :, None]
print(f'  warp sweep end: mean {w.mean():.0f}, earliest {w.min(axis=(1, 2)).mean():.0f}, latest {w.max(axis=(1, 2)).mean():.0f}; '
      f'spread inside a CTA (latest - earliest warp) mean {(w.max(axis=2) - w.min(axis=2)).mean():.0f}, max {(w.max(axis=2) - w.min(axis=2)).max(axis=1).mean():.0f}')
late = w.max(axis=2).mean(axis=0)
order = np.argsort(late)
print('  CTAs finishing their sweep last (mean ns):', [(int(c), int(late[c])) for c in order[-6:]], ' first:', [(int(c), int(late[c])) for c in order[:3]])
per_warp = w.mean(axis=(0, 1))
print('  mean sweep end by warp index:', np.round(per_warp).astype(int).tolist())
