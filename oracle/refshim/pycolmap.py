"""TEST INFRASTRUCTURE ONLY. Stand-in for the slice of pycolmap the reference uses
(/root/reference/sucre/sfm.py:188-223): Reconstruction(model_dir) reading a COLMAP *text* model, with
.cameras / .images dicts whose entries expose model.name, width, height, params, camera_id, image_id, name
and cam_from_world.rotation.matrix() / .translation.  The quaternion is normalised and converted with
Eigen's toRotationMatrix formula, in float64, like COLMAP does.
"""
from pathlib import Path
from types import SimpleNamespace

import numpy as np


class _Rotation:
    def __init__(self, q):
        q = np.asarray(q, dtype=np.float64)
        self._q = q / np.linalg.norm(q)

    def matrix(self):
        w, x, y, z = self._q
        tx, ty, tz = 2 * x, 2 * y, 2 * z
        twx, twy, twz = tx * w, ty * w, tz * w
        txx, txy, txz = tx * x, ty * x, tz * x
        tyy, tyz, tzz = ty * y, tz * y, tz * z
        return np.array([
            [1 - (tyy + tzz), txy - twz, txz + twy],
            [txy + twz, 1 - (txx + tzz), tyz - twx],
            [txz - twy, tyz + twx, 1 - (txx + tyy)],
        ])


class Reconstruction:
    def __init__(self, model_dir):
        model_dir = Path(model_dir)
        self.cameras = {}
        for line in (model_dir / 'cameras.txt').read_text().splitlines():
            if not line.strip() or line.startswith('#'):
                continue
            tok = line.split()
            cam_id = int(tok[0])
            self.cameras[cam_id] = SimpleNamespace(
                camera_id=cam_id, model=SimpleNamespace(name=tok[1]), width=int(tok[2]), height=int(tok[3]),
                params=np.array([float(x) for x in tok[4:]]))
        self.images = {}
        lines = [l for l in (model_dir / 'images.txt').read_text().splitlines() if not l.startswith('#')]
        for line in lines[0::2]:
            tok = line.split()
            if len(tok) < 10:
                continue
            image_id = int(tok[0])
            self.images[image_id] = SimpleNamespace(
                image_id=image_id, name=tok[9], camera_id=int(tok[8]),
                cam_from_world=SimpleNamespace(rotation=_Rotation([float(x) for x in tok[1:5]]),
                                               translation=np.array([float(x) for x in tok[5:8]])))
