"""TEST INFRASTRUCTURE ONLY. In-memory stand-in for the slice of h5py the reference uses
(/root/reference/sucre/loader.py:63-126): File(path, mode, libver=) as a context manager, iteration over
group names in NAME-SORTED order (h5py's default), create_group / create_dataset(name, data=), ds[()] get
and set, .shape, .name, .values(), .items().  Files live in a process-global dict keyed by path; a marker
file is touched on disk so that Path.exists()/unlink() in the reference behave.
"""
from pathlib import Path

import numpy as np

_STORE = {}


class _Dataset:
    def __init__(self, name, data):
        self.name = name
        self._data = np.array(data)

    @property
    def shape(self):
        return self._data.shape

    def __getitem__(self, key):
        assert key == ()
        return self._data.copy()

    def __setitem__(self, key, value):
        assert key == ()
        self._data[...] = value


class _Group:
    def __init__(self, name):
        self.name = name
        self._items = {}

    def create_dataset(self, name, data):
        self._items[name] = _Dataset(f'{self.name}/{name}', data)
        return self._items[name]

    def __getitem__(self, name):
        return self._items[name]


class File:
    def __init__(self, path, mode='r', libver=None):
        self._key = str(Path(path))
        if mode == 'r' and self._key not in _STORE:
            raise FileNotFoundError(self._key)
        if mode in ('a', 'r+'):
            _STORE.setdefault(self._key, {})
            Path(path).touch()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    @property
    def _groups(self):
        return _STORE[self._key]

    def create_group(self, name):
        if name in self._groups:
            raise ValueError(f'group {name} already exists')
        self._groups[name] = _Group('/' + name)
        return self._groups[name]

    def __getitem__(self, name):
        return self._groups[name]

    def __iter__(self):
        return iter(sorted(self._groups))

    def keys(self):
        return sorted(self._groups)

    def values(self):
        return [self._groups[k] for k in sorted(self._groups)]

    def items(self):
        return [(k, self._groups[k]) for k in sorted(self._groups)]


def _reset(path=None):
    if path is None:
        _STORE.clear()
    else:
        _STORE.pop(str(Path(path)), None)
