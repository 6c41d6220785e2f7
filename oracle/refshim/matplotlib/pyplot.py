"""TEST INFRASTRUCTURE ONLY (see package docstring).  `colormaps['jet']` is the one thing the reference touches
(plot_l, sucre.py:104): a callable mapping an (H,W) array in [0,1] to (H,W,4) RGBA floats.  A piecewise-linear
stand-in with jet's anchor colours is enough — the vignetting plot is not part of any parity check."""
import numpy as np


def _jet(x):
    x = np.clip(np.asarray(x, dtype=np.float64), 0.0, 1.0)
    r = np.clip(1.5 - np.abs(4 * x - 3), 0, 1)
    g = np.clip(1.5 - np.abs(4 * x - 2), 0, 1)
    b = np.clip(1.5 - np.abs(4 * x - 1), 0, 1)
    return np.stack([r, g, b, np.ones_like(x)], axis=-1)


colormaps = {'jet': _jet}
