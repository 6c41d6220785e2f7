#!/usr/bin/env python
"""Benchmark of the SUCRe hot path (gather + per-pixel fit) — contract in the task prompt / DESIGN.md §measurement.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (CUDA)
    python bench.py --impl reference [--gpus N] [--steps K] ...     # reference's CPU algorithm (oracle port)

One step = restore ONE target image of BASELINE.json configs[1] (synthetic 100-view 1368x912 scene): fused gather
against all views + 200 closed-form Adam iterations + final J.  `value` = pixel-views/s with the scene resident in
HBM; `e2e` = the same through api.restore_from_host (pinned host buffers, H2D + D2H inside the timed region; by
default only the footprint rectangle of every source view is copied, --upload full copies whole views).
N > 1: one process per GPU (torchrun), every rank restores a different target of its own replica of the scene
(weak scaling, no data-path collective); time = max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC, UNIT = 'pixel_views_per_s', 'pixel-views/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--views', type=int, default=100)
    ap.add_argument('--width', type=int, default=1368)
    ap.add_argument('--height', type=int, default=912)
    ap.add_argument('--num-iter', type=int, default=200)
    ap.add_argument('--target', type=int, default=None, help='target view index (default: a central view)')
    ap.add_argument('--cpu-sample-views', type=int, default=4)
    ap.add_argument('--cpu-sample-iters', type=int, default=2)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--upload', default='footprint', choices=['footprint', 'rows', 'full'],
                    help='what the end-to-end leg copies host -> device per step: whole views, or only the rectangle of '
                         'each source view the target can see (identical results, api.upload_plan)')
    ap.add_argument('--shard', default='targets', choices=['targets', 'pixels', 'pixels-nccl'],
                    help='N > 1: one target per rank (weak scaling, default) or ONE target sharded by pixel band over all '
                         'ranks (strong scaling; all-reduce fused into the fit kernel over NVLink, or through NCCL)')
    return ap.parse_args()


def workload(args) -> dict:
    return {'workload': f'synthetic {args.views}-view {args.width}x{args.height} PINHOLE scene, single target image '
                        f'per step, closed-form J, {args.num_iter} Adam iterations (BASELINE.json configs[1])',
            'views': args.views, 'width': args.width, 'height': args.height, 'num_iter': args.num_iter,
            'mode': 'use_closed_form', 'min_cover': 1e-6, 'seed': 0}


def default_target(args) -> int:
    import math
    grid = math.ceil(math.sqrt(args.views))
    return args.target if args.target is not None else min(args.views - 1, (grid // 2) * grid + grid // 2)


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, index: int):
        self.proc = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={index}', f'--query-gpu={self.FIELDS}',
                                          '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self) -> dict:
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        out, _ = self.proc.communicate(timeout=5)
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in out.splitlines():
            tok = [t.strip() for t in line.split(',')]
            if len(tok) < 7:
                continue
            try:
                sm.append(float(tok[0]))
                mx.append(float(tok[1]))
            except ValueError:
                continue
            for name, val in zip(names, tok[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


def measured_peak_gbs() -> tuple[float, str]:
    p = ROOT / 'MEASURED_PEAKS.json'
    if p.exists():
        try:
            return float(json.loads(p.read_text())['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def ncu_traffic_per_launch(n_obs: int):
    """dram bytes per fit_kernel launch from the committed ncu capture, if it was taken on this workload."""
    p = ROOT / 'profiles' / 'fit_kernel_traffic.json'
    if p.exists():
        t = json.loads(p.read_text())
        if abs(t.get('n_obs', -1) - n_obs) <= 0.02 * n_obs:
            return t['dram_bytes_per_launch']
    return None


# ---------------------------------------------------------------------------------------------------------------
def cpu_port_sample(args, target: int, render_device: str = 'cpu') -> dict:
    """Times the reference's CPU algorithm (oracle/torch_port.py: the same ATen ops as /root/reference/sucre) on a
    bounded sample of the workload and extrapolates linearly in views and iterations."""
    import torch
    from oracle import torch_port as tp
    from sucre_b200.synth import SyntheticScene
    sys.path.insert(0, str(ROOT / 'tests'))
    import helpers

    scene = SyntheticScene(args.views, args.width, args.height, seed=0)
    S = max(1, min(args.cpu_sample_views, args.views))
    stride = args.views // S
    sample = [stride // 2 + k * stride for k in range(S)]

    def view(i):
        K, R, t, W, H = helpers.reference_pose(scene, i)
        d, c = scene.render(i, device=render_device)
        return tp.make_view(K, R, t, W, H, d.cpu().to(torch.int32), c.cpu())

    tv = view(target)
    srcs = [(scene.image_name(i), view(i)) for i in sample]
    t0 = time.perf_counter()
    kept = tp.gather(tv, srcs, min_cover=1e-6)
    t_gather = time.perf_counter() - t0
    obs = [o for _, o in kept]
    n_obs = sum(o['u'].shape[0] for o in obs)
    model = tp.FormationModel(args.height, args.width, closed_form=True)
    t0 = time.perf_counter()
    tp.run_adam(model, obs, args.cpu_sample_iters, batch_size=5)
    t_fit = (time.perf_counter() - t0) / (args.cpu_sample_iters + 0.5)  # + the final closed-form J (half an iteration)
    scale = args.views / S
    t_step = (t_gather + t_fit * (args.num_iter + 0.5)) * scale
    return {'value': args.views * args.width * args.height / t_step, 'unit': UNIT, 'cores': torch.get_num_threads(),
            'kind': 'port',
            'sample': f'oracle/torch_port.py (reference ATen op sequence, in-RAM spill) on target {target} vs views {sample} '
                      f'({n_obs} observations): gather {t_gather:.2f} s, {args.cpu_sample_iters} closed-form Adam '
                      f'iterations at {t_fit:.2f} s/iter; extrapolated linearly to {args.views} views x {args.num_iter} '
                      f'iterations = {t_step:.0f} s per restored image',
            's_per_restored_image': t_step, 'gather_s': t_gather, 'fit_s_per_iter': t_fit, 'sample_obs': n_obs}


def reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    target = default_target(args)
    runs = []
    for i in range(args.warmup + args.steps):
        r = cpu_port_sample(args, target)
        if i >= args.warmup:
            runs.append(r)
        if i == 0 and args.warmup > 0 and r['gather_s'] + r['fit_s_per_iter'] * args.cpu_sample_iters > 60:
            break  # keep the whole arm within minutes on a slow host
    if not runs:
        runs = [r]
    best = min(runs, key=lambda x: x['s_per_restored_image'])
    mean_t = sum(x['s_per_restored_image'] for x in runs) / len(runs)
    value = args.views * args.width * args.height / mean_t
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': len(runs),
            'warmup': args.warmup, 'ms_per_step': mean_t * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': workload(args),
            's_per_restored_image': mean_t,
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': best['cores'], 'kind': 'port', 'sample': best['sample']},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
def ours(args):
    import torch
    import torch.distributed as dist
    from sucre_b200 import api, engine
    from sucre_b200.synth import SyntheticScene
    sys.path.insert(0, str(ROOT / 'tests'))
    import helpers

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)

    V, W, H = args.views, args.width, args.height
    scene = SyntheticScene(V, W, H, seed=0)
    geoms = []
    for i in range(V):
        K, R, t, w, h = helpers.reference_pose(scene, i)
        geoms.append(engine.ViewGeom.from_pose(K, R, t, w, h))
    depth, rgb = scene.render_all(device=dev)          # synthetic data, rendered straight into HBM
    host = api.HostScene(geoms, depth.cpu(), rgb.cpu()).pin()
    resident = engine.DeviceScene(dev)
    resident.add_views(list(range(V)), geoms, depth, rgb)
    target = (default_target(args) + rank) % V           # every rank restores a different target (weak scaling)
    keys = list(range(V))
    kw = dict(min_cover=1e-6, use_closed_form=True, num_iter=args.num_iter, lr=0.05)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    def run_steps(fn, n, fit_ms=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for _ in range(n):
            last = fn(fit_ms)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        barrier()
        if world > 1:
            tms = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)   # max over ranks, device-timed
            ms = float(tms.item())
        return ms, last

    def step_resident(fit_ms):
        if fit_ms is None:
            return api.restore_resident(resident, target, keys, **kw)
        # same call sequence as api.restore_resident, with events around the gather and the Adam loop (the dominant kernel)
        g0, f0, f1 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        g0.record()
        store = engine.gather(resident, target, keys, min_cover=1e-6)
        state = engine.FitState.initial(dev)
        f0.record()
        history = engine.fit(store, state, args.num_iter, 0.05)
        f1.record()
        J = engine.closed_form_J(store, state.params, state.J)
        fit_ms.append((f0, f1))
        gather_events.append((g0, f0))
        return api.RestoreResult(J=J, params=state.params, history=history, n_obs=store.n_obs, view_kept=store.view_kept)

    J_host = torch.empty((H, W, 3), dtype=torch.float32).pin_memory()

    def step_host(_):
        return api.restore_from_host(host, target, keys, device=dev, out_J=J_host, upload=args.upload, **kw)

    def step_host_full(_):
        return api.restore_from_host(host, target, keys, device=dev, out_J=J_host, upload='full', **kw)

    if args.shard != 'targets' and world > 1:
        return ours_pixel_sharded(args, resident, keys, dev, world, rank, local, run_steps, host)

    run_steps(step_resident, max(3, args.warmup))
    sampler = ClockSampler(local) if rank == 0 else None
    fit_events, gather_events = [], []
    ms_res, res = run_steps(step_resident, args.steps, fit_events)
    fit_ms = [a.elapsed_time(b) for a, b in fit_events]
    gather_ms = statistics.mean(a.elapsed_time(b) for a, b in gather_events)
    run_steps(step_host, 1)
    ms_e2e, res_h = run_steps(step_host, args.steps)
    clocks = sampler.stop() if sampler else None          # sampled across both timed regions (resident + end to end)
    # for comparison, outside the sampled regions: the same call copying whole views
    run_steps(step_host_full, 1)
    ms_full, res_f = run_steps(step_host_full, args.steps)

    pv_per_step = V * W * H * world                      # pixel-views all ranks process per step
    n_obs = res.n_obs
    value = pv_per_step / (ms_res / args.steps / 1e3)
    e2e_value = pv_per_step / (ms_e2e / args.steps / 1e3)
    peak, peak_src = measured_peak_gbs()
    fit_launch_us = statistics.mean(fit_ms) / args.num_iter * 1e3
    achieved = 16.0 * n_obs / (fit_launch_us * 1e-6) / 1e9   # algorithmic bytes per launch: 16 B per observation
    traffic = ncu_traffic_per_launch(n_obs)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': max(3, args.warmup),
        'ms_per_step': ms_res / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {**workload(args), 'parallelism': f'{world} rank(s), one target image per rank, scene replicated',
                   'target_rank0': target, 'observations_rank0': n_obs,
                   'l2': f'inputs larger than L2: scene {host.nbytes / 1e6:.0f} MB + observation store '
                         f'{16 * n_obs / 1e6:.0f} MB streamed every iteration (L2 126 MB)'},
        's_per_restored_image': ms_res / args.steps / 1e3,
        # SURVEY.md §8d(i) also quotes the gather stage alone: P*V / t_gather (match + plan + sample, its one host sync included)
        'gather': {'ms': gather_ms, 'pixel_views_per_s': V * W * H / (gather_ms / 1e3)},
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': res_h.h2d_bytes,
                'd2h_bytes_per_step': api.d2h_bytes(res_h), 's_per_restored_image': ms_e2e / args.steps / 1e3,
                'api': 'sucre_b200.api.restore_from_host (pinned host u16 depth + u8 colour in, J + parameters out)',
                'upload': f'{args.upload}: {res_h.h2d_bytes / 1e6:.0f} MB of the {host.nbytes / 1e6:.0f} MB host scene are copied '
                          f'per step (the target whole, of every source view the rectangle the target can see; same '
                          f'result bit for bit as --upload full, tests/test_upload_gpu.py)',
                'whole_view_upload': {'value': pv_per_step / (ms_full / args.steps / 1e3), 'unit': UNIT,
                                      'h2d_bytes_per_step': res_f.h2d_bytes,
                                      's_per_restored_image': ms_full / args.steps / 1e3}},
        'gpu_launches': args.steps * (api.LAUNCHES_FIXED + args.num_iter),
        'roofline': {'bound': 'hbm', 'kernel': 'fit_kernel<closed form>', 'achieved': achieved, 'peak': peak,
                     'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic, 'peak_source': peak_src,
                     'algorithmic_bytes_per_launch': 16 * n_obs, 'launch_us': fit_launch_us,
                     'launches_timed': len(fit_ms) * args.num_iter},
        'clocks': clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_port_sample(args, target, render_device=str(dev))
        line['cpu_baseline'] = {k: cpu[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def ours_pixel_sharded(args, resident, keys, dev, world, rank, local, run_steps, host):
    """Strong scaling: every step restores ONE target whose pixel bands are spread over all ranks (config 4 style)."""
    import torch
    import torch.distributed as dist
    from sucre_b200 import dist as sdist
    target = default_target(args)
    peers = sdist.PeerExchange(dev) if args.shard == 'pixels' else None

    def step(_):
        ops = sdist.CudaBandOps(resident, target, keys, use_closed_form=True)
        return sdist.restore_band_sharded(ops, min_cover=1e-6, num_iter=args.num_iter, lr=0.05, peers=peers)

    run_steps(step, max(3, args.warmup))
    sampler = ClockSampler(local) if rank == 0 else None
    ms, res = run_steps(step, args.steps)
    clocks = sampler.stop() if sampler else None
    if rank == 0:
        V, W, H = args.views, args.width, args.height
        print(json.dumps({
            'metric': METRIC, 'value': V * W * H / (ms / args.steps / 1e3), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(3, args.warmup), 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {**workload(args), 'parallelism': f'one target, pixel bands over {world} ranks, scene replicated; '
                       + ('all-reduce fused into fit_kernel over NVLink peer memory' if peers else 'NCCL all-reduce per iteration'),
                       'observations': res.n_obs},
            's_per_restored_image': ms / args.steps / 1e3, 'gpu_launches': args.steps * (api_launches(args) if peers else 0),
            'clocks': clocks}))
    dist.destroy_process_group()


def api_launches(args):
    from sucre_b200 import api
    return api.LAUNCHES_FIXED + args.num_iter


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        reference_arm(a)
    else:
        ours(a)
