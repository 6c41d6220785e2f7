"""Developer timing probe (not the bench): a synthetic scene of the given shape, CUDA events around each stage.
    python tools/quick_time.py [views width height [target [iters]]]"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
from sucre_b200 import engine, _lib  # noqa: E402
from sucre_b200.synth import SyntheticScene  # noqa: E402

V, W, H = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (100, 1368, 912)))
target = int(sys.argv[4]) if len(sys.argv) > 4 else V // 2 + 5
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 200
t0 = time.time()
scene = SyntheticScene(V, W, H, seed=0)
ds = engine.DeviceScene('cuda')
for i in range(V):
    K, R, t, w, h = scene.reference_pose(i)
    d, c = scene.render(i, device='cuda')
    ds.add_view(i, engine.ViewGeom.from_pose(K, R, t, w, h), d, c)
torch.cuda.synchronize()
print(f'scene built in {time.time()-t0:.1f}s')


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ts = []
    for _ in range(n):
        ev[0].record()
        r = fn()
        ev[1].record()
        torch.cuda.synchronize()
        ts.append(ev[0].elapsed_time(ev[1]))
    return min(ts), sum(ts) / len(ts), r


keys = list(range(V))
mn, av, store = timed(lambda: engine.gather(ds, target, keys))
print(f'gather: min {mn:.3f} ms avg {av:.3f} ms  N={store.n_obs} blocks={store.n_blocks} rows={store.n_rows} '
      f'kept={int(store.view_kept.sum())}/{V} obs/px={store.n_obs/(W*H):.2f} fill={store.fill:.3f} stats={store.stats} '
      f'-> {W*H*V/mn/1e6:.1f} G pixel-views/s')
L = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
trec = ds.record(target)
table = ds.table(tuple(keys))
nt = store.n_tiles
masks = torch.empty((nt, V), dtype=torch.int32, device='cuda')
import ctypes as C  # noqa: E402
band = store.band
mn, av, _ = timed(lambda: L.sucre_gather_match(trec.ctypes.data, table.data_ptr(), V, C.byref(band), masks.data_ptr(), 0, st))
print(f'  match kernel: min {mn:.3f} ms avg {av:.3f}')
pix = torch.empty(nt * 32, dtype=torch.int32, device='cuda')
pmasks = torch.empty_like(masks)
permuted = store.pix is not None
if permuted:
    mn, av, _ = timed(lambda: L.sucre_gather_permute(masks.data_ptr(), V, C.byref(band), W * H, pix.data_ptr(), pmasks.data_ptr(), st))
    print(f'  permute kernel: min {mn:.3f} ms avg {av:.3f}')
    masks = pmasks
vc = torch.empty(V, dtype=torch.int64, device='cuda')
vk = torch.empty(V, dtype=torch.uint8, device='cuda')
ro, bo, wo = (torch.empty(nt + 1, dtype=torch.int64, device='cuda') for _ in range(3))
tot = torch.empty(3, dtype=torch.int64, device='cuda')
mn, av, _ = timed(lambda: (L.sucre_gather_count(masks.data_ptr(), nt, V, vc.data_ptr(), st),
                           L.sucre_gather_plan(masks.data_ptr(), nt, V, vc.data_ptr(), W * H, 1e-6, vk.data_ptr(), ro.data_ptr(),
                                               bo.data_ptr(), wo.data_ptr(), tot.data_ptr(), st)))
print(f'  count + plan kernels: min {mn:.3f} ms avg {av:.3f}')
cells = torch.empty_like(store.cells)
bm = torch.empty(max(1, store.n_blocks), dtype=torch.int32, device='cuda')
bv = torch.empty_like(bm)
mn, av, _ = timed(lambda: L.sucre_gather_sample(trec.ctypes.data, table.data_ptr(), V, C.byref(band), pix.data_ptr() if permuted else 0,
                                                masks.data_ptr(), vk.data_ptr(),
                                                wo.data_ptr(), bo.data_ptr(), store.record_format, cells.data_ptr(),
                                                bm.data_ptr(), bv.data_ptr(), 0, st))
print(f'  sample kernel: min {mn:.3f} ms avg {av:.3f}  ({store.stream_bytes/mn/1e6:.0f} GB/s of rows written)')


def fit():
    s = engine.FitState.initial('cuda')
    h = engine.fit(store, s, iters)
    return s, h


mn, av, (s, h) = timed(fit, 3)
print(f'fit {iters} it: min {mn:.3f} ms avg {av:.3f} ms -> {mn/iters*1e3:.1f} us/iter, '
      f'{store.n_obs*store.record_bytes*iters/mn/1e6:.0f} GB/s algorithmic ({store.record_bytes} B/obs), '
      f'{store.stream_bytes*iters/mn/1e6:.0f} GB/s streamed')
print('params', s.params.cpu().numpy(), 'cost', h[0, 9].item(), h[-1, 9].item())
mn, av, J = timed(lambda: engine.closed_form_J(store, s.params))
print(f'write_J: min {mn:.3f} ms')
