"""Developer probe: where a fit_kernel launch spends its tail.  Needs a library built with
`make -C sucre_b200/csrc clean all EXTRA=-DSUCRE_FIT_TIMING` (the kernel then leaves %globaltimer stamps in the fit
workspace).  Prints, for the LAST of `iters` chained launches: spread of the CTA start times, of the warp end times
within and across CTAs, and the serial tail (last warp end -> end of the last CTA's reduction + Adam step)."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / 'tests')]
import helpers  # noqa: E402
from sucre_b200 import engine  # noqa: E402
from sucre_b200.synth import SyntheticScene  # noqa: E402

V, W, H = 100, 1368, 912
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 50
scene = SyntheticScene(V, W, H, seed=0)
ds = engine.DeviceScene('cuda')
depth, rgb = scene.render_all(device='cuda')
ds.add_views(list(range(V)), [engine.ViewGeom.from_pose(*helpers.reference_pose(scene, i)) for i in range(V)], depth, rgb)
store = engine.gather(ds, 55, list(range(V)))
state = engine.FitState.initial('cuda')
engine.fit(store, state, iters)
torch.cuda.synchronize()
MAXC, WARPS = 2048, 16
ws = store.workspace.cpu().numpy()
base = ws.size - 8 * (MAXC + MAXC * WARPS + 1)
t = ws[base:].view(np.uint64).astype(np.int64)
n = 148
start, wend, last = t[:n], t[MAXC:MAXC + n * WARPS].reshape(n, WARPS), t[MAXC + MAXC * WARPS]
t0 = start.min()
cta_end = wend.max(axis=1)
print(f'CTA starts: spread {(start.max() - t0) / 1e3:.2f} us')
print(f'warp ends rel. first start (us): min {(wend.min() - t0) / 1e3:.2f} mean {(wend.mean() - t0) / 1e3:.2f} max {(wend.max() - t0) / 1e3:.2f}')
print(f'CTA ends (slowest warp): min {(cta_end.min() - t0) / 1e3:.2f} mean {(cta_end.mean() - t0) / 1e3:.2f} max {(cta_end.max() - t0) / 1e3:.2f}')
print(f'within-CTA spread (max - min warp end): mean {(wend.max(1) - wend.min(1)).mean() / 1e3:.2f} max {(wend.max(1) - wend.min(1)).max() / 1e3:.2f} us')
print(f'CTA busy time (end - own start): min {(cta_end - start).min() / 1e3:.2f} mean {(cta_end - start).mean() / 1e3:.2f} max {(cta_end - start).max() / 1e3:.2f} us')
print(f'serial tail after the last warp: {(last - wend.max()) / 1e3:.2f} us; launch total {(last - t0) / 1e3:.2f} us')
order = np.argsort(cta_end)
print('slowest CTAs', order[-5:].tolist(), 'fastest', order[:5].tolist())

# ---- per-tile work features + the partition, for offline calibration of the cost model (gpurun_out/fit_timing.npz)
if len(sys.argv) > 2:
    dev = store.cells.device
    G = store.seg_views
    nblk_tile = store.blk_off[1:] - store.blk_off[:-1]
    blk_tile = torch.repeat_interleave(torch.arange(store.n_tiles, device=dev), nblk_tile)
    j = torch.arange(store.n_blocks, device=dev) - store.blk_off[blk_tile]
    seg = store.seg_off[blk_tile] + j // G
    lanes = torch.arange(32, device=dev, dtype=torch.int64)
    bits = ((store.blk_mask.to(torch.int64)[:, None] >> lanes[None, :]) & 1).to(torch.int32)
    cnt = torch.zeros((store.n_segments, 32), dtype=torch.int32, device=dev).index_add_(0, seg, bits)
    seg_tile = torch.repeat_interleave(torch.arange(store.n_tiles, device=dev), store.seg_off[1:] - store.seg_off[:-1])

    def per_tile(x):
        return torch.zeros(store.n_tiles, dtype=torch.int64, device=dev).index_add_(0, seg_tile, x.to(torch.int64)).cpu().numpy()

    feats = dict(blocks=nblk_tile.cpu().numpy(), segments=(store.seg_off[1:] - store.seg_off[:-1]).cpu().numpy(),
                 records=(store.rec_off[1:] - store.rec_off[:-1]).cpu().numpy(),
                 steps=per_tile(cnt.max(1).values), pairs=per_tile((cnt // 2).max(1).values),
                 odd=per_tile(((cnt % 2) > 0).any(1)), active=per_tile((cnt > 0).sum(1)))
    part_off = 8 * 10 * MAXC
    partition = ws[part_off:part_off + 4 * (n * WARPS + 1)].view(np.int32)
    Path(sys.argv[2]).parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(sys.argv[2], start=start, wend=wend, last=last, partition=partition, **feats)
    print('saved', sys.argv[2])
