"""End-to-end time of api.restore_from_host per upload mode (BASELINE configs[1] shape by default):
    python tools/upload_modes.py [views width height num_iter]"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / 'tests')]
import helpers  # noqa: E402
from sucre_b200 import api, engine  # noqa: E402
from sucre_b200.synth import SyntheticScene  # noqa: E402

V, W, H, iters = (int(x) for x in (sys.argv[1:5] + ['100', '1368', '912', '200'][len(sys.argv) - 1:]))
dev = torch.device('cuda', 0)
scene = SyntheticScene(V, W, H, seed=0)
geoms = [engine.ViewGeom.from_pose(*helpers.reference_pose(scene, i)) for i in range(V)]
depth, rgb = scene.render_all(device=dev)
host = api.HostScene(geoms, depth.cpu(), rgb.cpu()).pin()
del depth, rgb
grid = int(-(-V ** 0.5 // 1))
target = min(V - 1, (grid // 2) * grid + grid // 2)
J_host = torch.empty((H, W, 3), dtype=torch.float32).pin_memory()
out = {}
for mode in ('full', 'rows', 'footprint', 'full', 'footprint'):
    for steps in (2, 5):  # warm-up, then timed
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            r = api.restore_from_host(host, target, list(range(V)), device=dev, out_J=J_host, upload=mode,
                                      use_closed_form=True, num_iter=iters)
        e1.record()
        torch.cuda.synchronize()
    out.setdefault(mode, []).append({'ms_per_image': e0.elapsed_time(e1) / steps, 'h2d_MB': r.h2d_bytes / 1e6, 'n_obs': r.n_obs})
print(json.dumps(out))
