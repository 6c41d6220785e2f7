"""CPU-only: the reference arm of bench.py prints one JSON line with the contract's keys; rank != 0 prints nothing."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
SMALL = ['--views', '6', '--width', '96', '--height', '64', '--num-iter', '5', '--steps', '1', '--warmup', '0']


def _run(env=None):
    e = dict(os.environ, **(env or {}))
    return subprocess.run([sys.executable, str(ROOT / 'bench.py'), '--impl', 'reference', '--gpus', '2', *SMALL],
                          capture_output=True, text=True, env=e, timeout=300)


def test_reference_arm_line():
    r = _run()
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'pixel_views_per_s' and d['unit'] == 'pixel-views/s'
    assert d['higher_is_better'] is True and d['vs_baseline'] is None and d['n_gpus'] == 2 and d['steps'] == 1
    assert d['value'] > 0 and abs(d['value'] - 6 * 96 * 64 / (d['ms_per_step'] / 1e3)) < 1e-6 * d['value']
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and 'torch_port' in d['cpu_baseline']['sample']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['config']['workload'].startswith('synthetic 6-view 96x64')
    assert set(d['config']) == {'workload', 'views', 'width', 'height', 'num_iter', 'mode', 'min_cover', 'seed'}   # same keys in both arms
    m = d['measured']
    assert len(m['iteration_s']) == 1 and 'ALL 6 views' in d['cpu_baseline']['sample']
    s_image = m['gather_s'] + 5 * m['iteration_s'][0] + m['final_J_s']   # extrapolated in iterations only
    assert abs(d['s_per_restored_image'] - s_image) < 1e-9 * s_image


def test_reference_arm_other_ranks_exit_quietly():
    r = _run({'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
    assert r.returncode == 0 and r.stdout.strip() == ''
