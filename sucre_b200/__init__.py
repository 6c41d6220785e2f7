"""sucre_b200 — B200-native implementation of SUCRe's data-parallel hot path.

Two stages, both hand-written sm_100a CUDA behind a C ABI (include/sucre_b200.h):
  * multi-view correspondence gather  (reference: sucre/sfm.py:90-138, 154-175; sucre/loader.py:78-87, 103-118)
  * per-pixel fit of the underwater image formation model (reference: sucre/sucre.py:52-82, 124-157)
The Python modules mirror the reference's call surface (sfm.COLMAPModel, sucre.restore_image, the CLI).
There is no CPU fallback: every op raises if the CUDA library is missing.
"""
__version__ = '0.1.0'
