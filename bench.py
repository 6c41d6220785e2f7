#!/usr/bin/env python
"""Benchmark of the SUCRe hot path (gather + per-pixel fit) — contract in the task prompt / DESIGN.md §6.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (CUDA)
    python bench.py --impl reference [--gpus N] [--steps K] ...     # the reference's CPU algorithm (oracle port)

One step = restore ONE target image of BASELINE.json configs[1] (synthetic 100-view 1368x912 scene): fused gather
against all views + 200 closed-form Adam iterations + final J.  `value` = pixel-views/s with the scene resident in
HBM; `e2e` = the same through api.restore_stream (pinned host buffers, H2D + D2H of every step inside the timed
region, double-buffered so that the copies of neighbouring steps overlap the fit; by default only the footprint
rectangle of every source view is copied, --upload full copies whole views).

N > 1 (one process per GPU, torchrun): the SAME single target is restored by all N ranks — each rank gathers and fits
a band of its pixels, the per-iteration all-reduce of the 10 residual sums runs inside the fit kernel over NVLink
peer memory, the J bands are written straight into rank 0's buffer (`scaling: "strong"`).  In the same run rank 0
also restores the target alone and the two results are compared (`parity`); a mismatch beyond 1e-5 exits non-zero.
Extra keys: `target_parallel` (one target per rank, no collective: the weak-scaling figure) and, at N = 8 (or with
--config4 on), `config4` (200 views 3840x2160, one target over all ranks, with its own single-GPU time and parity).

Reference arm: the reference's own ATen op sequence (oracle/torch_port.py, pinned bit for bit to the unmodified
reference by tests/test_torch_port.py; the reference itself is Python that does not travel to the GPU box) on ALL
views of the same scene, on the host cores: full gather + sampling timed once, then W + K full-size closed-form Adam
iterations (each step = one iteration over all observations, the last K timed), then the final update_J.  The
per-image time is extrapolated in ITERATIONS ONLY: gather + num_iter x mean(iteration) + final J.
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
PARITY_TOL = 1e-5


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
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-budget-s', type=float, default=480.0,
                    help='reference arm: stop timing iterations once this much wall time went into them (>= 2 are always timed)')
    ap.add_argument('--upload', default='footprint', choices=['footprint', 'rows', 'full'],
                    help='what the end-to-end leg copies host -> device per step: whole views, or only the rectangle of '
                         'each source view the target can see (identical results, api.upload_plan)')
    ap.add_argument('--shard', default='pixels', choices=['pixels', 'pixels-nccl', 'targets'],
                    help='N > 1: ONE target sharded by pixel band over all ranks (strong scaling; all-reduce fused into the '
                         'fit kernel over NVLink, or through NCCL), or one target per rank (weak scaling, no collective)')
    ap.add_argument('--config4', default='auto', choices=['auto', 'on', 'off'],
                    help='N > 1: also time the config-4 shape (200 views 3840x2160, one target over all ranks); auto = at N = 8')
    return ap.parse_args()


def workload(args) -> dict:
    """Identical in both arms (the driver compares it)."""
    return {'workload': f'synthetic {args.views}-view {args.width}x{args.height} PINHOLE scene, single target image '
                        f'per step, closed-form J, {args.num_iter} Adam iterations (BASELINE.json configs[1])',
            'views': args.views, 'width': args.width, 'height': args.height, 'num_iter': args.num_iter,
            'mode': 'use_closed_form', 'min_cover': 1e-6, 'seed': 0}


def default_target(views: int, target=None) -> int:
    import math
    grid = math.ceil(math.sqrt(views))
    return target if target is not None else min(views - 1, (grid // 2) * grid + grid // 2)


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
        if abs(t.get('n_obs', -1) - n_obs) <= 0.02 * n_obs and t.get('record_bytes') == 8:
            return t['dram_bytes_per_launch']
    return None


# ---------------------------------------------------------------------------------------------------------------
def reference_arm(args):
    """The reference's CPU algorithm on the host cores (see the module docstring).  Must not initialise CUDA."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    from oracle import torch_port as tp
    from sucre_b200.synth import SyntheticScene

    V, W, H = args.views, args.width, args.height
    target = default_target(V, args.target)
    scene = SyntheticScene(V, W, H, seed=0)
    t_setup = time.perf_counter()

    def view(i):
        K, R, t, w, h = scene.reference_pose(i)
        d, c = scene.render(i, device='cpu')
        return tp.make_view(K, R, t, w, h, d.to(torch.int32), c)

    tv = view(target)
    srcs = [(scene.image_name(i), tv if i == target else view(i)) for i in range(V)]
    t_setup = time.perf_counter() - t_setup

    t0 = time.perf_counter()
    kept = tp.gather(tv, srcs, min_cover=1e-6)                    # sfm.py:127-138 + loader.py:78-87, 103-118, all views
    t_gather = time.perf_counter() - t0
    obs = [o for _, o in kept]
    n_obs = sum(o['u'].shape[0] for o in obs)
    model = tp.FormationModel(H, W, closed_form=True)
    opt = torch.optim.Adam(model.parameters(), lr=0.05)
    for _ in range(args.warmup):
        tp.adam_iteration(model, obs, opt, n_obs, batch_size=5)
    t_iter = []
    for _ in range(max(1, args.steps)):
        t0 = time.perf_counter()
        tp.adam_iteration(model, obs, opt, n_obs, batch_size=5)   # sucre.py:138-148, all observations
        t_iter.append(time.perf_counter() - t0)
        if len(t_iter) >= 2 and sum(t_iter) > args.cpu_budget_s:
            break
    t0 = time.perf_counter()
    model.solve_J(obs)                                            # sucre.py:156
    t_final = time.perf_counter() - t0
    mean_iter = statistics.mean(t_iter)
    s_image = t_gather + args.num_iter * mean_iter + t_final
    value = V * W * H / s_image
    sample = (f'oracle/torch_port.py (reference ATen op sequence, in-RAM spill) on {cores} host threads, target {target} vs ALL {V} views: '
              f'gather + sampling {t_gather:.2f} s ({len(kept)} views kept, {n_obs} observations), {len(t_iter)} full-size closed-form '
              f'Adam iterations at {mean_iter:.2f} s/iter (min {min(t_iter):.2f}, max {max(t_iter):.2f}; {args.warmup} untimed before), '
              f'final update_J {t_final:.2f} s; extrapolated in iterations only: gather + {args.num_iter} x iteration + final J = '
              f'{s_image:.0f} s per restored image')
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': len(t_iter),
            'warmup': args.warmup, 'ms_per_step': s_image * 1e3, 'higher_is_better': True, 'scaling': 'strong' if args.gpus > 1 else 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': workload(args),
            's_per_restored_image': s_image,
            'step_definition': 'one full-size closed-form Adam iteration of the reference algorithm over all observations; '
                               'ms_per_step is the per-image time extrapolated in iterations only',
            'measured': {'gather_s': t_gather, 'iteration_s': t_iter, 'final_J_s': t_final, 'setup_s': t_setup,
                         'timed_wall_s': t_gather + sum(t_iter) + t_final, 'observations': n_obs, 'views_kept': len(kept)},
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


def cpu_baseline_subprocess(args) -> dict | None:
    """`cpu_baseline` of our arm: the reference arm in a FRESH process that never sees a GPU, before this process
    touches CUDA (2 timed full-size iterations after 1 warm-up)."""
    cmd = [sys.executable, str(ROOT / 'bench.py'), '--impl', 'reference', '--steps', '2', '--warmup', '1',
           '--views', str(args.views), '--width', str(args.width), '--height', str(args.height),
           '--num-iter', str(args.num_iter)]
    if args.target is not None:
        cmd += ['--target', str(args.target)]
    env = {k: v for k, v in os.environ.items() if k not in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK')}
    env['CUDA_VISIBLE_DEVICES'] = ''
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=1500)
        line = [l for l in r.stdout.splitlines() if l.startswith('{')][-1]
        d = json.loads(line)
        return {**d['cpu_baseline'], 's_per_restored_image': d['s_per_restored_image'], 'measured': d['measured']}
    except Exception as e:  # the bench line still carries everything else
        return {'value': None, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'port', 'sample': f'failed: {e!r}'}


# ---------------------------------------------------------------------------------------------------------------
def build_scene(V, W, H, dev):
    """Synthetic scene rendered straight into HBM + its geometry."""
    from sucre_b200 import engine
    from sucre_b200.synth import SyntheticScene
    scene = SyntheticScene(V, W, H, seed=0)
    geoms = [engine.ViewGeom.from_pose(*scene.reference_pose(i)) for i in range(V)]
    depth, rgb = scene.render_all(device=dev)
    resident = engine.DeviceScene(dev)
    resident.add_views(list(range(V)), geoms, depth, rgb)
    return geoms, depth, rgb, resident


def gather_roofline(store, V, W, H, gather_ms, peak):
    """SURVEY.md §8d's algorithmic bytes of the gather, restated for 8-byte records:
    2 P (target depth) + per view [2 n_inbounds (source depth) + 3 n (source colour) + 8 n (record) + P/8 (mask bits)]."""
    P, N, n_inb = W * H, store.n_obs, store.stats['n_inbounds']
    alg = 2 * P + 2 * n_inb + (3 + store.record_bytes) * N + V * P // 8
    gbs = alg / (gather_ms * 1e-3) / 1e9
    return {'ms': gather_ms, 'pixel_views_per_s': V * P / (gather_ms / 1e3), 'n_inbounds': n_inb, 'n_obs': N,
            'tile_views_culled': store.stats['tile_views_culled'], 'tile_views': store.stats['tile_views'],
            'pixel_views_culled': store.stats['tile_views_culled'] * 32,
            'algorithmic_bytes': alg, 'achieved': gbs, 'peak': peak, 'unit': 'GB/s', 'frac': gbs / peak,
            'formula': '2P + 2 n_inbounds + (3 + record_bytes) N + V P / 8 (SURVEY.md 8d with 8-byte records); time = match + '
                       'count + plan + sample + the one host sync that sizes the store'}


def ours(args):
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_subprocess(args)     # before this process initialises CUDA

    import torch
    import torch.distributed as dist
    from sucre_b200 import api, engine

    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)

    V, W, H = args.views, args.width, args.height
    geoms, depth, rgb, resident = build_scene(V, W, H, dev)
    host = api.HostScene(geoms, depth.cpu(), rgb.cpu()).pin()
    keys = list(range(V))
    kw = dict(min_cover=1e-6, use_closed_form=True, num_iter=args.num_iter, lr=0.05)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    def run_steps(fn, n, *a):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for _ in range(n):
            last = fn(*a)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        barrier()
        if world > 1:
            tms = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)   # max over ranks, device-timed
            ms = float(tms.item())
        return ms, last

    ctx = dict(args=args, world=world, rank=rank, local=local, dev=dev, resident=resident, host=host, keys=keys, kw=kw,
               run_steps=run_steps, barrier=barrier, geoms=geoms)
    if world > 1:
        if args.shard == 'targets':
            line = ours_target_parallel(ctx)
        else:
            line = ours_pixel_sharded(ctx)
        if rank == 0:
            print(json.dumps(line), flush=True)
        bad = torch.tensor([1 if rank == 0 and not line.get('parity_ok', True) else 0], device=dev)
        dist.all_reduce(bad)     # every rank leaves with rank 0's verdict
        barrier()
        dist.destroy_process_group()
        sys.exit(3 if int(bad.item()) else 0)

    # ---- one GPU ------------------------------------------------------------------------------------------------
    target = default_target(V, args.target)

    def step_resident(events):
        if events is None:
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
        events.append((g0, f0, f1))
        return api.RestoreResult(J=J, params=state.params, history=history, n_obs=store.n_obs, view_kept=store.view_kept,
                                 store=store)

    J_pair = [torch.empty((H, W, 3), dtype=torch.float32).pin_memory() for _ in range(2)]

    def stream_host(n, upload):
        """n targets back to back through the double-buffered host pipeline: every step copies its inputs H2D and its
        results D2H inside the timed region; the copies of step k+1 / k-1 overlap the fit of step k."""
        last = None
        for last in api.restore_stream(host, [target] * n, keys, device=dev, upload=upload, out_J=J_pair, **kw):
            pass
        return last

    def single_host(upload):
        return api.restore_from_host(host, target, keys, device=dev, out_J=J_pair[0], upload=upload, **kw)

    warm = max(3, args.warmup)
    run_steps(step_resident, warm, None)
    sampler = ClockSampler(local)
    events = []
    ms_res, res = run_steps(step_resident, args.steps, events)
    fit_ms = [f0.elapsed_time(f1) for _, f0, f1 in events]
    gather_ms = statistics.mean(g0.elapsed_time(f0) for g0, f0, _ in events)
    run_steps(stream_host, 1, 3, args.upload)
    ms_e2e, res_h = run_steps(stream_host, 1, args.steps, args.upload)
    clocks = sampler.stop()                              # sampled across both timed regions (resident + end to end)
    # for comparison, outside the sampled regions: one un-pipelined call per step, and the pipeline copying whole views
    run_steps(single_host, 1, args.upload)
    ms_single, _ = run_steps(single_host, args.steps, args.upload)
    run_steps(stream_host, 1, 2, 'full')
    ms_full, res_f = run_steps(stream_host, 1, args.steps, 'full')

    pv_per_step = V * W * H
    store = res.store
    n_obs = res.n_obs
    value = pv_per_step / (ms_res / args.steps / 1e3)
    e2e_value = pv_per_step / (ms_e2e / args.steps / 1e3)
    peak, peak_src = measured_peak_gbs()
    fit_launch_us = statistics.mean(fit_ms) / args.num_iter * 1e3
    alg_bytes = store.record_bytes * n_obs                   # each observation read once: 8 B = {z f32, colour u8 x 3, pad}
    achieved = alg_bytes / (fit_launch_us * 1e-6) / 1e9
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': 1, 'steps': args.steps, 'warmup': warm,
        'ms_per_step': ms_res / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic', 'config': workload(args),
        'parallelism': '1 rank, scene resident in HBM', 'target': target, 'observations': n_obs,
        'l2': f'inputs larger than L2: scene {host.nbytes / 1e6:.0f} MB + observation store {store.stream_bytes / 1e6:.0f} MB '
              f'streamed every iteration (L2 126 MB)',
        's_per_restored_image': ms_res / args.steps / 1e3,
        # SURVEY.md §8d(i) also quotes the gather stage alone: P*V / t_gather
        'roofline_gather': gather_roofline(store, V, W, H, gather_ms, peak),
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': res_h.h2d_bytes,
                'd2h_bytes_per_step': api.d2h_bytes(res_h), 's_per_restored_image': ms_e2e / args.steps / 1e3,
                'api': 'sucre_b200.api.restore_stream (pinned host u16 depth + u8 colour in, J + parameters + history out; '
                       'double-buffered: the H2D of step k+1 and the D2H of step k-1 overlap the fit of step k, every step '
                       'copies its own inputs and results)',
                'upload': f'{args.upload}: {res_h.h2d_bytes / 1e6:.0f} MB of the {host.nbytes / 1e6:.0f} MB host scene are copied '
                          f'per step (the target whole, of every source view the rectangle the target can see; same '
                          f'result bit for bit as --upload full, tests/test_upload_gpu.py)',
                'single_call': {'s_per_restored_image': ms_single / args.steps / 1e3,
                                'api': 'sucre_b200.api.restore_from_host, one blocking call per step (no overlap across steps)'},
                'whole_view_upload': {'value': pv_per_step / (ms_full / args.steps / 1e3), 'unit': UNIT,
                                      'h2d_bytes_per_step': res_f.h2d_bytes,
                                      's_per_restored_image': ms_full / args.steps / 1e3}},
        'gpu_launches': args.steps * api.LAUNCHES_PER_IMAGE,
        'gpu_launches_note': 'per image: gather_match, count_views, permute, tile_count, scan, gather_sample, partition, ONE resident fit_kernel '
                             f'launch running all {args.num_iter} Adam iterations, fit_kernel<write J>',
        'roofline': {'bound': 'hbm', 'kernel': 'fit_kernel<closed form, 8-byte records>', 'achieved': achieved, 'peak': peak,
                     'unit': 'GB/s', 'frac': achieved / peak, 'traffic': ncu_traffic_per_launch(n_obs), 'peak_source': peak_src,
                     'algorithmic_bytes_per_launch': alg_bytes, 'launch_us': fit_launch_us,
                     'launches_timed': len(fit_ms) * args.num_iter,
                     'record': '8 B per observation {z f32, colour u8 x 3, pad}: SURVEY.md 8d restated (the reference\'s I is '
                               'exactly u8 / 255, loader.py:157); with its 16-byte float payload the same launch is '
                               f'{16 * n_obs / (fit_launch_us * 1e-6) / 1e9:.0f} GB/s',
                     'streamed_bytes_per_launch': store.stream_bytes, 'fill': store.fill},
        'clocks': clocks,
    }
    if cpu is not None:
        line['cpu_baseline'] = cpu
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
def parity_report(sharded, single) -> dict:
    """Sharded result against the single-GPU restore of the same target (rank 0)."""
    import torch
    p1, pN = single.params.double(), sharded.params.double()
    rel = float(((pN - p1).abs() / p1.abs()).max())
    J1, JN = single.J.reshape(-1, 3), sharded.J.reshape(-1, 3)
    nan_eq = bool(torch.equal(torch.isnan(J1), torch.isnan(JN)))
    j_abs = float((JN - J1).nan_to_num(0.0).abs().max())
    n_eq = int(sharded.n_obs) == int(single.n_obs)
    return {'params_rel': rel, 'J_maxabs': j_abs, 'nan_set_equal': nan_eq, 'n_obs_equal': n_eq, 'tolerance': PARITY_TOL,
            'ok': bool(rel <= PARITY_TOL and j_abs <= PARITY_TOL and nan_eq and n_eq)}


def sharded_block(ctx, resident, keys, target, num_iter, steps, warm, peers, sdist, with_e2e_host=None):
    """Times ONE target restored by all ranks (resident scene) and checks it against rank 0's single-GPU restore."""
    import torch
    from sucre_b200 import api
    rank, dev, run_steps = ctx['rank'], ctx['dev'], ctx['run_steps']

    def step():
        ops = sdist.CudaBandOps(resident, target, keys, use_closed_form=True)
        return sdist.restore_band_sharded(ops, min_cover=1e-6, num_iter=num_iter, lr=0.05, peers=peers, root_only=True)

    run_steps(step, warm)
    ms, res = run_steps(step, steps)
    J_sharded = res.J.clone() if rank == 0 else None
    status = int(res.status.item()) if res.status is not None else 0
    # the same target on rank 0 alone (the others wait at the barrier): single-GPU time + parity
    ms1, parity = None, None
    if rank == 0:
        one = lambda: api.restore_resident(resident, target, keys, min_cover=1e-6, use_closed_form=True, num_iter=num_iter, lr=0.05)  # noqa: E731
        for _ in range(3):   # the whole-image store is new to this rank's allocator: let it settle before timing
            one()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n1 = max(3, steps)
        e0.record()
        for _ in range(n1):
            single = one()
        e1.record()
        torch.cuda.synchronize(dev)
        ms1 = e0.elapsed_time(e1) / n1
        res.J = J_sharded
        parity = parity_report(res, single)
        parity['exchange_status'] = status
        parity['ok'] = bool(parity['ok'] and status == 0)
    ctx['barrier']()
    return ms / steps, ms1, parity, res


def ours_pixel_sharded(ctx):
    """Strong scaling: every step restores ONE target whose pixel bands are spread over all ranks."""
    import torch
    from sucre_b200 import api
    from sucre_b200 import dist as sdist
    args, world, rank, local, dev = ctx['args'], ctx['world'], ctx['rank'], ctx['local'], ctx['dev']
    V, W, H = args.views, args.width, args.height
    target = default_target(V, args.target)
    peers = sdist.PeerExchange(dev) if args.shard == 'pixels' else None
    warm = max(3, args.warmup)
    sampler = ClockSampler(local) if rank == 0 else None
    ms_step, ms_single, parity, res = sharded_block(ctx, ctx['resident'], ctx['keys'], target, args.num_iter, args.steps, warm, peers, sdist)
    clocks = sampler.stop() if sampler else None

    # end to end: every rank copies what ITS band needs from pinned host memory, restores, rank 0 reads J back; the
    # copies of step k+1 / k-1 overlap the fit of step k (api.restore_stream_sharded); one blocking call per step beside it
    J_pair = [torch.empty((H, W, 3), dtype=torch.float32).pin_memory() for _ in range(2)] if rank == 0 else None

    def stream_host(n):
        last = None
        for last in api.restore_stream_sharded(ctx['host'], [target] * n, ctx['keys'], device=dev, peers=peers, out_J=J_pair,
                                               upload=args.upload, **ctx['kw']):
            pass
        return last

    def step_host():
        return api.restore_from_host_sharded(ctx['host'], target, ctx['keys'], device=dev, peers=peers,
                                             out_J=None if J_pair is None else J_pair[0], upload=args.upload, **ctx['kw'])
    ctx['run_steps'](stream_host, 1, 3)
    ms_e2e, res_h = ctx['run_steps'](stream_host, 1, args.steps)
    ctx['run_steps'](step_host, 2)
    ms_single_call, _ = ctx['run_steps'](step_host, args.steps)
    h2d = torch.tensor([res_h.h2d_bytes], dtype=torch.int64, device=dev)
    import torch.distributed as dist
    dist.all_reduce(h2d)

    # one target per rank, no collective (the weak-scaling figure of round 1), for context
    tp = None
    t_rank = (target + rank) % V
    ms_tp, _ = ctx['run_steps'](lambda: api.restore_resident(ctx['resident'], t_rank, ctx['keys'], **ctx['kw']), args.steps)
    tp = {'value': V * W * H * world / (ms_tp / args.steps / 1e3), 'unit': UNIT, 'ms_per_step': ms_tp / args.steps,
          'scaling': 'weak', 'note': 'every rank restores its own target of its replica of the scene; no data-path collective'}

    line = {
        'metric': METRIC, 'value': V * W * H / (ms_step / 1e3), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': warm, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic', 'config': workload(args),
        'parallelism': f'ONE target, pixel bands over {world} ranks, scene replicated; '
                       + ('per-iteration all-reduce of the 10 sums fused into fit_kernel over NVLink peer memory, J bands written into rank 0 over NVLink'
                          if peers else 'NCCL all-reduce per iteration, NCCL all-gather of J'),
        'target': target, 'observations': res.n_obs, 'observations_rank0': res.n_local,
        's_per_restored_image': ms_step / 1e3,
        'single_gpu_same_run': {'ms_per_step': ms_single, 'speedup': None if ms_single is None else ms_single / ms_step,
                                'efficiency': None if ms_single is None else ms_single / ms_step / world},
        'overhead_us_per_iteration_vs_ideal': None if ms_single is None else (ms_step - ms_single / world) / args.num_iter * 1e3,
        'parity': parity,
        'e2e': {'value': V * W * H / (ms_e2e / args.steps / 1e3), 'unit': UNIT, 'h2d_bytes_per_step': int(h2d.item()),
                'd2h_bytes_per_step': H * W * 3 * 4 + 9 * 4 + args.num_iter * 10 * 4, 's_per_restored_image': ms_e2e / args.steps / 1e3,
                'api': 'sucre_b200.api.restore_stream_sharded (every rank uploads the rectangles its band can see; J + parameters '
                       'read back on rank 0; double-buffered: the copies of step k+1 / k-1 overlap the fit of step k)',
                'single_call': {'s_per_restored_image': ms_single_call / args.steps / 1e3,
                                'api': 'sucre_b200.api.restore_from_host_sharded, one blocking call per step (no overlap across steps)'}},
        'gpu_launches': args.steps * world * api.LAUNCHES_PER_BAND,
        'target_parallel': tp,
        'clocks': clocks,
    }
    want4 = args.config4 == 'on' or (args.config4 == 'auto' and world == 8)
    if want4:
        line['config4'] = config4_block(ctx, peers, sdist)
    if rank == 0:
        line['parity_ok'] = bool(parity['ok'] and (not want4 or line['config4']['parity']['ok']))
    return line


def config4_block(ctx, peers, sdist, V=200, W=3840, H=2160):
    """BASELINE.json configs[3]: synthetic 200-view 4K scene, single target, pixel-tile sharded fit over all ranks."""
    import torch
    world, rank, dev = ctx['world'], ctx['rank'], ctx['dev']
    torch.cuda.empty_cache()
    geoms, depth, rgb, resident = build_scene(V, W, H, dev)
    del depth, rgb
    keys = list(range(V))
    target = default_target(V)
    steps, warm = 3, 2
    ms_step, ms_single, parity, res = sharded_block(ctx, resident, keys, target, ctx['args'].num_iter, steps, warm, peers, sdist)
    return {'workload': f'synthetic {V}-view {W}x{H} scene, single target, closed-form J, {ctx["args"].num_iter} Adam iterations '
                        f'(BASELINE.json configs[3]), pixel bands over {world} ranks',
            'value': V * W * H / (ms_step / 1e3), 'unit': UNIT, 'ms_per_step': ms_step, 'steps': steps, 'warmup': warm,
            'observations': res.n_obs, 'target': target,
            'single_gpu_same_run': {'ms_per_step': ms_single, 'speedup': None if ms_single is None else ms_single / ms_step,
                                    'efficiency': None if ms_single is None else ms_single / ms_step / world},
            'overhead_us_per_iteration_vs_ideal': None if ms_single is None else (ms_step - ms_single / world) / ctx['args'].num_iter * 1e3,
            'parity': parity}


def ours_target_parallel(ctx):
    """--shard targets: one target per rank, no data-path collective (weak scaling)."""
    from sucre_b200 import api
    args, world, rank, local = ctx['args'], ctx['world'], ctx['rank'], ctx['local']
    V, W, H = args.views, args.width, args.height
    target = (default_target(V, args.target) + rank) % V
    step = lambda: api.restore_resident(ctx['resident'], target, ctx['keys'], **ctx['kw'])  # noqa: E731
    warm = max(3, args.warmup)
    ctx['run_steps'](step, warm)
    sampler = ClockSampler(local) if rank == 0 else None
    ms, res = ctx['run_steps'](step, args.steps)
    clocks = sampler.stop() if sampler else None
    return {'metric': METRIC, 'value': V * W * H * world / (ms / args.steps / 1e3), 'unit': UNIT, 'n_gpus': world,
            'steps': args.steps, 'warmup': warm, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': workload(args),
            'parallelism': f'{world} ranks, one target image per rank, scene replicated, no data-path collective',
            'observations_rank0': res.n_obs, 's_per_restored_image': ms / args.steps / 1e3 / world,
            'gpu_launches': args.steps * world * api.LAUNCHES_PER_IMAGE, 'clocks': clocks}


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        reference_arm(a)
    else:
        ours(a)
