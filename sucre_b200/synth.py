"""Seeded synthetic underwater survey: the only data source of tests and bench.

An analytic sea floor is imaged by nadir PINHOLE cameras on a lawn-mower grid.  Every view gets a u16
millimetre depth map (with one 8x8 invalid patch) and a u8 RGB image produced by pushing an analytic
texture through the underwater image formation model with known ground-truth parameters.  All geometry is
computed in float64 and then quantised, so the integer outputs do not depend on the device or on the SIMD
flavour of sin/cos (a code flips only if a float64 value lies within 1 ulp of a rounding boundary).

Nothing here is on the hot path: the scene stands in for the PNG files + COLMAP model the reference reads
(/root/reference/sucre/loader.py:156-170, /root/reference/sucre/sfm.py:186-226).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np
import torch

GT_B = (0.05, 0.25, 0.35)
GT_BETA = (0.45, 0.12, 0.08)
GT_GAMMA = (0.35, 0.20, 0.15)


@dataclass
class SyntheticScene:
    """Host-side description of a synthetic scene; pixels are rendered lazily, view by view."""
    n_views: int
    width: int
    height: int
    seed: int = 0
    # optional second camera (width, height): every `alt_every`-th view (i % alt_every == alt_every-1) uses it
    alt_size: tuple[int, int] | None = None
    alt_every: int = 0
    cams: list = field(init=False, repr=False)      # [(W, H, fx, fy, cx, cy)], COLMAP camera ids are index+1
    view_cam: np.ndarray = field(init=False, repr=False)  # (V,) index into cams
    # world_from_cam rotation (V,3,3) and camera centre (V,3), float64
    R_wc: np.ndarray = field(init=False, repr=False)
    C: np.ndarray = field(init=False, repr=False)
    patch: np.ndarray = field(init=False, repr=False)  # (V,2) top-left corner (u,v) of the invalid 8x8 patch

    def __post_init__(self):
        rng = np.random.default_rng(self.seed)
        V = self.n_views
        self.cams = [(self.width, self.height, 0.9 * self.width, 0.9 * self.width, self.width / 2, self.height / 2)]
        self.view_cam = np.zeros(V, dtype=np.int64)
        if self.alt_size is not None and self.alt_every > 0:
            aw, ah = self.alt_size
            self.cams.append((aw, ah, 0.8 * aw, 0.85 * aw, aw / 2 + 1.25, ah / 2 - 0.75))
            self.view_cam[np.arange(V) % self.alt_every == self.alt_every - 1] = 1
        grid = math.ceil(math.sqrt(V))
        idx = np.arange(V)
        row = idx // grid
        col = idx % grid
        col = np.where(row % 2 == 1, grid - 1 - col, col)  # lawn-mower: odd rows run backwards
        C = np.stack([col * 0.35, row * 0.30, np.full(V, 2.0)], axis=1).astype(np.float64)
        C[:, :2] += rng.normal(0.0, 0.03, size=(V, 2))
        C[:, 2] += rng.normal(0.0, 0.1, size=V)
        rpy = rng.normal(0.0, 0.05, size=(V, 3))
        nadir = np.diag([1.0, -1.0, -1.0])  # camera z looks down, camera y points to world -y
        R = np.empty((V, 3, 3))
        for i in range(V):
            R[i] = _rot_z(rpy[i, 2]) @ _rot_y(rpy[i, 1]) @ _rot_x(rpy[i, 0]) @ nadir
        self.R_wc = R
        self.C = C
        wh = np.array([self.cams[k][:2] for k in self.view_cam])
        self.patch = np.stack([rng.integers(0, wh[:, 0] - 8), rng.integers(0, wh[:, 1] - 8)], axis=1)

    # -- naming -------------------------------------------------------------------------------------
    def image_name(self, i: int) -> str:
        return f'image{i:04d}.png'

    def depth_name(self, i: int) -> str:
        return f'depth_image{i:04d}.png'

    # -- COLMAP convention: cam_from_world quaternion (w,x,y,z) + translation ------------------------
    def cam_from_world(self, i: int) -> tuple[np.ndarray, np.ndarray]:
        R_cw = self.R_wc[i].T
        t_cw = -R_cw @ self.C[i]
        return _quat_from_rot(R_cw), t_cw

    def reference_pose(self, i: int):
        """K, R, t (cam->world) as the reference's COLMAPModel derives them (sfm.py:204-208, 219-222), from the
        COLMAP cam_from_world quaternion + translation (quaternion -> matrix in float64 with Eigen's formula).
        Returns (K (3,3), R (3,3), t (3,1), width, height), float32 torch tensors."""
        from .sfm import quaternion_to_matrix
        q, t_cw = self.cam_from_world(i)
        R_cw = torch.tensor(quaternion_to_matrix(q), dtype=torch.float32)
        t_cw = torch.tensor(t_cw, dtype=torch.float32).view(3, 1)
        W, H, fx, fy, cx, cy = self.cams[self.view_cam[i]]
        K = torch.tensor([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=torch.float32)
        return K, R_cw.T, -R_cw.T @ t_cw, W, H

    # -- rendering ----------------------------------------------------------------------------------
    @torch.no_grad()
    def render(self, i: int, device: str | torch.device = 'cpu') -> tuple[torch.Tensor, torch.Tensor]:
        """Returns (depth u16 (H,W) millimetres, rgb u8 (H,W,3)) of view `i` as tensors on `device`."""
        f64 = dict(dtype=torch.float64, device=device)
        W, H, fx, fy, cx, cy = self.cams[self.view_cam[i]]
        u = torch.arange(W, **f64) + 0.5
        v = torch.arange(H, **f64) + 0.5
        dx = ((u - cx) / fx)[None, :].expand(H, W)
        dy = ((v - cy) / fy)[:, None].expand(H, W)
        R = torch.tensor(self.R_wc[i], **f64)
        C = torch.tensor(self.C[i], **f64)
        # world ray direction of the camera ray (dx, dy, 1); s is then the camera-z depth along it
        wx = R[0, 0] * dx + R[0, 1] * dy + R[0, 2]
        wy = R[1, 0] * dx + R[1, 1] * dy + R[1, 2]
        wz = R[2, 0] * dx + R[2, 1] * dy + R[2, 2]
        s = (0.0 - C[2]) / wz
        for _ in range(30):  # fixed point of C_z + s*wz = h(C_x + s*wx, C_y + s*wy)
            s = (_height(C[0] + s * wx, C[1] + s * wy) - C[2]) / wz
        x = C[0] + s * wx
        y = C[1] + s * wy
        depth_mm = torch.round(s * 1000.0).clamp_(0, 65535)
        pu, pv = int(self.patch[i, 0]), int(self.patch[i, 1])
        depth_mm[pv:pv + 8, pu:pu + 8] = 0
        rng_m = s * torch.sqrt(dx * dx + dy * dy + 1.0)  # metric range camera -> sea floor
        rgb = torch.empty((H, W, 3), **f64)
        for c in range(3):
            J = _texture(x, y, c)
            rgb[..., c] = J * torch.exp(-GT_BETA[c] * rng_m) + GT_B[c] * (1.0 - torch.exp(-GT_GAMMA[c] * rng_m))
        rgb_u8 = torch.round(rgb.clamp_(0.0, 1.0) * 255.0).to(torch.uint8)
        return depth_mm.to(torch.int32).to(torch.uint16), rgb_u8

    def render_all(self, device: str | torch.device = 'cpu', views=None) -> tuple[torch.Tensor, torch.Tensor]:
        """Stacked (V,H,W) u16 depth and (V,H,W,3) u8 rgb for `views` (default: all)."""
        assert len(self.cams) == 1, 'render_all needs a single camera'
        views = list(range(self.n_views)) if views is None else list(views)
        depth = torch.empty((len(views), self.height, self.width), dtype=torch.uint16, device=device)
        rgb = torch.empty((len(views), self.height, self.width, 3), dtype=torch.uint8, device=device)
        for k, i in enumerate(views):
            depth[k], rgb[k] = self.render(i, device=device)
        return depth, rgb

    # -- on-disk form (PNG + COLMAP text model), what the reference CLI consumes ---------------------
    def write(self, root: Path, views=None, binary_model: bool = False) -> dict[str, Path]:
        import cv2
        root = Path(root)
        dirs = {k: root / k for k in ('images', 'depth', 'model')}
        for d in dirs.values():
            d.mkdir(parents=True, exist_ok=True)
        views = list(range(self.n_views)) if views is None else list(views)
        for i in views:
            depth, rgb = self.render(i)
            cv2.imwrite(str(dirs['depth'] / self.depth_name(i)), depth.numpy())
            cv2.imwrite(str(dirs['images'] / self.image_name(i)), np.ascontiguousarray(rgb.numpy()[..., ::-1]))
        write_colmap_text(self, dirs['model'], views)
        if binary_model:
            write_colmap_binary(self, dirs['model'], views)
        return dirs


def write_colmap_text(scene: SyntheticScene, model_dir: Path, views=None):
    views = list(range(scene.n_views)) if views is None else list(views)
    with open(model_dir / 'cameras.txt', 'w') as f:
        f.write('# Camera list with one line of data per camera:\n#   CAMERA_ID, MODEL, WIDTH, HEIGHT, PARAMS[]\n')
        for k, (W, H, fx, fy, cx, cy) in enumerate(scene.cams):
            f.write(f'{k + 1} PINHOLE {W} {H} {fx!r} {fy!r} {cx!r} {cy!r}\n')
    with open(model_dir / 'images.txt', 'w') as f:
        f.write('# Image list with two lines of data per image:\n'
                '#   IMAGE_ID, QW, QX, QY, QZ, TX, TY, TZ, CAMERA_ID, NAME\n#   POINTS2D[] as (X, Y, POINT3D_ID)\n')
        for i in views:
            q, t = scene.cam_from_world(i)
            f.write(' '.join([str(i + 1)] + [repr(float(x)) for x in (*q, *t)]
                             + [str(scene.view_cam[i] + 1), scene.image_name(i)]) + '\n\n')
    with open(model_dir / 'points3D.txt', 'w') as f:
        f.write('# 3D point list (empty)\n')


def write_colmap_binary(scene: SyntheticScene, model_dir: Path, views=None):
    import struct
    views = list(range(scene.n_views)) if views is None else list(views)
    with open(model_dir / 'cameras.bin', 'wb') as f:
        f.write(struct.pack('<Q', len(scene.cams)))
        for k, (W, H, fx, fy, cx, cy) in enumerate(scene.cams):
            f.write(struct.pack('<iiQQ', k + 1, 1, W, H))  # model id 1 = PINHOLE
            f.write(struct.pack('<4d', fx, fy, cx, cy))
    with open(model_dir / 'images.bin', 'wb') as f:
        f.write(struct.pack('<Q', len(views)))
        for i in views:
            q, t = scene.cam_from_world(i)
            f.write(struct.pack('<i7di', i + 1, *q, *t, int(scene.view_cam[i]) + 1))
            f.write(scene.image_name(i).encode() + b'\x00')
            f.write(struct.pack('<Q', 0))
    with open(model_dir / 'points3D.bin', 'wb') as f:
        f.write(struct.pack('<Q', 0))


# ---------------------------------------------------------------------------------------------------
def _height(x, y):
    return 0.15 * torch.sin(1.3 * x) * torch.cos(0.9 * y) + 0.05 * torch.sin(4.1 * x + 2.2 * y)


def _texture(x, y, c: int):
    """Smooth analytic albedo in [0.1, 0.9], different per channel."""
    a = (2.1, 1.7, 2.6)[c]
    b = (1.9, 2.3, 1.4)[c]
    p = (0.3, 1.1, 2.0)[c]
    return 0.5 + 0.25 * torch.sin(a * x + p) * torch.cos(b * y - p) + 0.15 * torch.sin(5.3 * x - 3.7 * y + 2 * p)


def _rot_x(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=np.float64)


def _rot_y(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float64)


def _rot_z(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float64)


def _quat_from_rot(R: np.ndarray) -> np.ndarray:
    """Unit quaternion (w,x,y,z) of a rotation matrix (Shepperd's method)."""
    tr = np.trace(R)
    if tr > 0:
        s = math.sqrt(tr + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = math.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s])
    elif R[1, 1] > R[2, 2]:
        s = math.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        q = np.array([(R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s])
    else:
        s = math.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        q = np.array([(R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s])
    return q / np.linalg.norm(q)
