"""Developer probe: how evenly do the rows of a config-2 store split over N ranks — contiguous bands of tiles
(dist.tile_band) against a cyclic assignment of 64-tile chunks."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT)]
from sucre_b200 import engine  # noqa: E402
from sucre_b200.dist import tile_band  # noqa: E402
from sucre_b200.synth import SyntheticScene  # noqa: E402

V, W, H = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (100, 1368, 912)))
target = int(sys.argv[4]) if len(sys.argv) > 4 else 55
scene = SyntheticScene(V, W, H, seed=0)
ds = engine.DeviceScene('cuda')
for i in range(V):
    d, c = scene.render(i, device='cuda')
    ds.add_view(i, engine.ViewGeom.from_pose(*scene.reference_pose(i)), d, c)
store = engine.gather(ds, target, list(range(V)))
rows = (store.row_off[1:] - store.row_off[:-1]).cpu().numpy()
T = len(rows)
print(f'{T} tiles, {rows.sum()} rows, rows/tile mean {rows.mean():.1f} min {rows.min()} max {rows.max()}')
for n in (2, 4, 8):
    band = [rows[lo:lo + m].sum() for lo, m in (tile_band(T, r, n) for r in range(n))]
    chunk = 64
    cyc = [sum(rows[c * chunk:(c + 1) * chunk].sum() for c in range(r, (T + chunk - 1) // chunk, n)) for r in range(n)]
    print(f'N={n}: contiguous bands max/mean = {max(band) / np.mean(band):.4f}   cyclic 64-tile chunks max/mean = {max(cyc) / np.mean(cyc):.4f}')
