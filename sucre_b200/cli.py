"""Command line of the restoration, flag-compatible with the reference's `python sucre.py ...`
(reference: sucre/sucre.py:222-307): same option names, defaults, mutual exclusion and target / pairing selection.

    python -m sucre_b200.sucre --image-dir D --depth-dir D --model-dir D --output-dir D (--image-name N | --image-list F | --image-ids A B) [...]

Multi-GPU (configs 3 and 5): launched under torchrun, one process per GPU,

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 -m sucre_b200.sucre ... --image-ids 1 1001

every rank restores a contiguous share of the targets on its own GPU (cuda:LOCAL_RANK), decoding only the views its
targets overlap.  Targets are independent (sucre.py:243-261 is a plain loop), so no process group is created and
nothing is exchanged; the output files are named after the targets, so ranks never write the same file.
"""
from __future__ import annotations

import argparse
import os
from pathlib import Path

# (flag, argparse keywords) in the reference's order; defaults are the reference's (sucre.py:265-305)
_DIRECTORIES = [
    ('--image-dir', 'directory of the undistorted colour images'),
    ('--depth-dir', 'directory of the 16-bit millimetre depth maps (depth_<image stem>.png)'),
    ('--model-dir', 'undistorted COLMAP model (cameras/images as .bin or .txt), PINHOLE cameras only'),
    ('--output-dir', 'where <stem>.pt, <stem>_rgb.png and <stem>_reconstruction.png are written'),
]
_OPTIONS = [
    ('--light-model', dict(action='store_true', help='also fit the artificial-light cone (10 extra parameters)')),
    ('--use-closed-form', dict(action='store_true',
                               help='J follows in closed form from B, beta, gamma every iteration instead of being an '
                                    'Adam parameter itself')),
    ('--min-cover', dict(type=float, default=0.000001,
                         help='a view is paired only if its matches cover more than this fraction of the target')),
    ('--image-scale', dict(type=float, default=1.0, help='work on images rescaled by this factor')),
    ('--filter-images-path', dict(type=Path, help='text file of image names (one per line) never used as source views')),
    ('--learning-rate', dict(type=float, default=0.05, help='Adam step size')),
    ('--num-iter', dict(type=int, default=200, help='Adam iterations')),
    ('--batch-size', dict(type=int, default=5,
                          help='accepted for compatibility: the CUDA fit streams every observation each iteration, so '
                               'view batching has no effect')),
    ('--save-interval', dict(type=int, help='also save the plots every this many iterations')),
    ('--params-path', dict(type=Path, help='.pt file to warm-start the model parameters from')),
    ('--force-compute-matches', dict(action='store_true', help='recompute matches even if a kept matches file exists')),
    ('--keep-matches', dict(action='store_true', help='keep the matches of every target on disk (large)')),
    ('--num-workers', dict(type=int, default=0, help='decode threads (0 = decode in the main thread)')),
    ('--device', dict(type=str, default='cuda', help='CUDA device of the computation')),
]


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(description='SUCRe.', formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    for flag, text in _DIRECTORIES:
        parser.add_argument(flag, required=True, type=Path, help=text)
    which = parser.add_mutually_exclusive_group(required=True)
    which.add_argument('--image-name', type=str, help='restore this image')
    which.add_argument('--image-list', type=Path, help='restore the images named in this text file, one per line')
    which.add_argument('--image-ids', type=int, nargs=2, metavar=('MIN_ID', 'MAX_ID'),
                       help='restore the images whose COLMAP ids lie in [MIN_ID, MAX_ID)')
    for flag, kw in _OPTIONS:
        parser.add_argument(flag, **kw)
    return parser


def parse_args(args: argparse.Namespace):
    """Builds the model, resolves targets and pairing list (sucre.py:222-261) and restores every target."""
    from . import loader, sfm, sucre
    print('Loading COLMAP model.')
    colmap_model = sfm.COLMAPModel(model_dir=args.model_dir, image_dir=args.image_dir, depth_dir=args.depth_dir,
                                   image_scale=args.image_scale)
    if args.image_name is not None:
        targets = [colmap_model[args.image_name]]
    elif args.image_list is not None:
        targets = [colmap_model[name] for name in args.image_list.read_text().splitlines()]
    else:  # ids missing from the model are skipped
        targets = [colmap_model.images[i] for i in range(*args.image_ids) if i in colmap_model.images]

    # under torchrun: this rank's share of the targets, on this rank's GPU
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (('WORLD_SIZE', '1'), ('RANK', '0'), ('LOCAL_RANK', '0')))
    if world > 1:
        from .dist import shard_targets
        targets = shard_targets(targets, rank, world, contiguous=True)
        if args.device == 'cuda':
            args.device = f'cuda:{local}'
        if args.num_workers > 0:   # the ranks share the host's cores: keep the decode pools from oversubscribing them
            args.num_workers = max(1, min(args.num_workers, (os.cpu_count() or 1) // world))
        print(f'Rank {rank}/{world}: {len(targets)} target(s) on {args.device}.')

    excluded = set(args.filter_images_path.read_text().splitlines()) if args.filter_images_path else set()
    pairing = [im for im in colmap_model.images.values() if im.name not in excluded]

    args.output_dir.mkdir(parents=True, exist_ok=True)
    # several targets: their output files are written by background threads while the next target is restored
    with loader.AsyncWriter() as writer:
        for image in targets:
            sucre.restore_image(
                image=image, colmap_model=colmap_model, output_dir=args.output_dir, light_model=args.light_model,
                use_closed_form=args.use_closed_form, min_cover=args.min_cover, image_list=pairing,
                lr=args.learning_rate, num_iter=args.num_iter, batch_size=args.batch_size,
                save_interval=args.save_interval, params_path=args.params_path,
                force_compute_matches=args.force_compute_matches, keep_matches=args.keep_matches,
                num_workers=args.num_workers, device=args.device, **({'writer': writer} if len(targets) > 1 else {}))


def main(argv=None):
    parse_args(build_parser().parse_args(argv))


if __name__ == '__main__':
    main()
