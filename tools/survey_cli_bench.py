"""End-to-end, WITH FILES, of the multi-target configs (BASELINE.json configs[2] and configs[4]) through the CLI:

    python tools/survey_cli_bench.py --views 300 --targets 50 --gpus 1 2 4 8        # config 3: --image-list of 50 targets
    python tools/survey_cli_bench.py --views 1000 --targets all --gpus 8            # config 5: --image-ids over the survey

Writes the synthetic survey to disk once (lossless PNG colour + 16-bit PNG depth + COLMAP text model, what the
reference CLI consumes), then for every N runs

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        -m sucre_b200.sucre --image-dir ... --depth-dir ... --model-dir ... --output-dir ... --use-closed-form (--image-list F | --image-ids A B)

(N = 1: plain `python -m sucre_b200.sucre`) and reports the wall time of the whole job — process start, COLMAP model,
PNG decode, gather + fit, .pt and PNG outputs — as seconds per restored image and pixel-views/s.  One JSON line per N
on stdout; with --out also appended to that file.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def write_survey(root: Path, V: int, W: int, H: int):
    """Renders on the GPU when there is one (identical pixels: the scene is quantised from float64) and encodes PNGs on
    a thread pool."""
    import cv2
    import numpy as np
    import torch
    from sucre_b200.synth import SyntheticScene, write_colmap_text
    scene = SyntheticScene(V, W, H, seed=0)
    dirs = {k: root / k for k in ('images', 'depth', 'model')}
    for d in dirs.values():
        d.mkdir(parents=True, exist_ok=True)
    dev = 'cuda' if torch.cuda.is_available() else 'cpu'

    def encode(i, depth, rgb):
        cv2.imwrite(str(dirs['depth'] / scene.depth_name(i)), depth)
        cv2.imwrite(str(dirs['images'] / scene.image_name(i)), np.ascontiguousarray(rgb[..., ::-1]))

    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as pool:
        jobs = []
        for i in range(V):
            depth, rgb = scene.render(i, device=dev)
            jobs.append(pool.submit(encode, i, depth.cpu().numpy(), rgb.cpu().numpy()))
        for j in jobs:
            j.result()
    write_colmap_text(scene, dirs['model'])
    return scene, dirs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--views', type=int, default=300)
    ap.add_argument('--width', type=int, default=1368)
    ap.add_argument('--height', type=int, default=912)
    ap.add_argument('--targets', default='50', help='number of targets (an evenly spread --image-list), or "all" (--image-ids over the survey)')
    ap.add_argument('--gpus', type=int, nargs='+', default=[1])
    ap.add_argument('--num-iter', type=int, default=200)
    ap.add_argument('--out', type=Path, default=None)
    ap.add_argument('--keep', type=Path, default=None, help='write the survey here and keep it (default: a temporary directory)')
    ap.add_argument('--profile', type=Path, default=None,
                    help='also run the 1-GPU job under cProfile and write the 40 most expensive calls (cumulative) here')
    args = ap.parse_args()
    V, W, H = args.views, args.width, args.height
    tmp = Path(tempfile.mkdtemp(prefix='sucre_survey_')) if args.keep is None else args.keep
    try:
        t0 = time.time()
        scene, dirs = write_survey(tmp, V, W, H)
        t_write = time.time() - t0
        if args.targets == 'all':
            n_targets, sel = V, ['--image-ids', '1', str(V + 1)]
        else:
            n_targets = int(args.targets)
            idx = [round(k * (V - 1) / max(1, n_targets - 1)) for k in range(n_targets)]
            (tmp / 'targets.txt').write_text('\n'.join(scene.image_name(i) for i in idx) + '\n')
            sel = ['--image-list', str(tmp / 'targets.txt')]
        for n in args.gpus:
            out_dir = tmp / f'out_{n}'
            shutil.rmtree(out_dir, ignore_errors=True)
            cli = ['-m', 'sucre_b200.sucre', '--image-dir', str(dirs['images']), '--depth-dir', str(dirs['depth']),
                   '--model-dir', str(dirs['model']), '--output-dir', str(out_dir), '--use-closed-form',
                   '--num-iter', str(args.num_iter), '--num-workers', '8', *sel]
            launcher = [sys.executable] if n == 1 else \
                [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={n}', '--master-addr', '127.0.0.1',
                 '--master-port', str(29500 + n)]
            env = dict(os.environ, PYTHONPATH=str(ROOT))
            t0 = time.time()
            r = subprocess.run(launcher + cli, cwd=ROOT, env=env, capture_output=True, text=True)
            wall = time.time() - t0
            n_pt = len(list(out_dir.glob('*.pt')))
            n_png = len(list(out_dir.glob('*_rgb.png')))
            line = {'config': f'synthetic {V}-view {W}x{H} survey, {n_targets} targets ({sel[0]}), closed-form J, {args.num_iter} '
                              f'Adam iterations, PNG files in, .pt + PNG files out',
                    'n_gpus': n, 'wall_s': wall, 's_per_restored_image': wall / n_targets,
                    'pixel_views_per_s': n_targets * V * W * H / wall, 'outputs': {'pt': n_pt, 'rgb_png': n_png},
                    'ok': r.returncode == 0 and n_pt == n_targets and n_png == n_targets, 'survey_write_s': t_write,
                    'launch': ' '.join(launcher[1:] + cli[:2]) + ' ...'}
            if not line['ok']:
                line['stderr_tail'] = r.stderr[-1500:]
            print(json.dumps(line), flush=True)
            if args.out is not None:
                args.out.parent.mkdir(parents=True, exist_ok=True)
                with open(args.out, 'a') as f:
                    f.write(json.dumps(line) + '\n')
        if args.profile is not None:
            out_dir, stats = tmp / 'out_profile', tmp / 'cli.prof'
            cli = ['-m', 'sucre_b200.sucre', '--image-dir', str(dirs['images']), '--depth-dir', str(dirs['depth']),
                   '--model-dir', str(dirs['model']), '--output-dir', str(out_dir), '--use-closed-form',
                   '--num-iter', str(args.num_iter), '--num-workers', '8', *sel]
            t0 = time.time()
            subprocess.run([sys.executable, '-m', 'cProfile', '-o', str(stats)] + cli, cwd=ROOT, env=dict(os.environ, PYTHONPATH=str(ROOT)),
                           capture_output=True, text=True)
            wall = time.time() - t0
            import io
            import pstats
            buf = io.StringIO()
            pstats.Stats(str(stats), stream=buf).sort_stats('cumulative').print_stats(40)
            args.profile.parent.mkdir(parents=True, exist_ok=True)
            args.profile.write_text(f'wall {wall:.2f} s under cProfile (main thread only)\n' + buf.getvalue())
    finally:
        if args.keep is None:
            shutil.rmtree(tmp, ignore_errors=True)


if __name__ == '__main__':
    main()
