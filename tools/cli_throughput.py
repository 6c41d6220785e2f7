"""End-to-end CLI throughput with files: PNG scene on disk -> `python -m sucre_b200.sucre --image-list ...` ->
.pt / PNG outputs.  Reports wall seconds per restored image (decode amortised over the targets)."""
import sys, time, tempfile
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from sucre_b200 import sucre
from sucre_b200.synth import SyntheticScene

V, W, H, T = (int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (30, 1368, 912, 10)))
with tempfile.TemporaryDirectory() as tmp:
    tmp = Path(tmp)
    t0 = time.time()
    scene = SyntheticScene(V, W, H, seed=0)
    dirs = scene.write(tmp)
    print(f'wrote {V} views in {time.time() - t0:.1f} s')
    names = [scene.image_name(i) for i in range(V // 2 - T // 2, V // 2 - T // 2 + T)]
    (tmp / 'targets.txt').write_text('\n'.join(names) + '\n')
    argv = ['--image-dir', str(dirs['images']), '--depth-dir', str(dirs['depth']), '--model-dir', str(dirs['model']),
            '--output-dir', str(tmp / 'out'), '--image-list', str(tmp / 'targets.txt'), '--use-closed-form']
    import io, contextlib
    for label in ('cold (incl. CUDA init, decode of all views)', 'warm process (decode again, kernels warm)'):
        t0 = time.time()
        with contextlib.redirect_stdout(io.StringIO()):
            sucre.main(argv)
        dt = time.time() - t0
        print(f'{label}: {T} targets in {dt:.2f} s -> {dt / T * 1e3:.0f} ms per restored image with file outputs')
    if len(sys.argv) > 5:
        import cProfile, pstats
        pr = cProfile.Profile()
        pr.enable()
        with contextlib.redirect_stdout(io.StringIO()):
            sucre.main(argv)
        pr.disable()
        pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
