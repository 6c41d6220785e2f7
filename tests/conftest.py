import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / 'tests' / 'golden'


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


class Golden:
    """A fixture file written by oracle/gen_golden.py from the unmodified reference."""

    def __init__(self, name: str):
        self.z = np.load(GOLDEN / f'{name}.npz', allow_pickle=False)
        self.name = name

    def __getitem__(self, k):
        return self.z[k]

    def __contains__(self, k):
        return k in self.z.files

    @property
    def n_views(self):
        return int(self.z['n_views'])

    def view_index(self, image_name: str) -> int:
        return int(str(image_name)[len('image'):len('image') + 4])

    def geom_arrays(self, i: int) -> dict:
        return {k: self.z[f'ref_{k}_{i}'] for k in ('K', 'Kinv', 'R', 't', 'Ri', 'ti', 'wh')}

    def inputs(self, i: int):
        return self.z[f'in_depth_{i}'], self.z[f'in_rgb_{i}']

    def matches(self, image_name: str) -> dict:
        return {k: self.z[f'm_{image_name}_{k}'] for k in ('u1', 'v1', 'u2', 'v2', 'd', 'I', 'cP', 'z')}

    def pairing_list(self):
        """View indices the reference matched against (all views minus the filtered ones), in model order."""
        filtered = set(self.z['filtered'].tolist()) if 'filtered' in self.z.files else set()
        return [i for i, n in enumerate(self.z['names'].tolist()) if n not in filtered]


@pytest.fixture(scope='session')
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = Golden(name)
        return cache[name]
    return get
