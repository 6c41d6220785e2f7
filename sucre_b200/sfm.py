"""Geometry data model with the reference's call surface (reference: sucre/sfm.py).

Same names and meaning as the reference — Pose, Camera, Image, COLMAPModel, Image.match_images — but:
  * the COLMAP model is parsed here (text or binary), pycolmap is not needed (sfm.py:186-226 used it);
  * images are decoded once into a device-resident scene (u16 depth, u8 colour) instead of being re-decoded for
    every (target, view) pair (sfm.py:130-133);
  * Image.match_images enqueues the fused CUDA gather and leaves its result on the device, in the
    MatchesFile's observation store, instead of appending to an HDF5 file (loader.py:68-76).
"""
from __future__ import annotations

import os
import struct
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np
import torch
from torch import Tensor

from . import loader
from .engine import DeviceScene, ViewGeom, gather

# COLMAP camera model id -> (name, number of parameters)
_CAMERA_MODELS = {0: ('SIMPLE_PINHOLE', 3), 1: ('PINHOLE', 4), 2: ('SIMPLE_RADIAL', 4), 3: ('RADIAL', 5),
                  4: ('OPENCV', 8), 5: ('OPENCV_FISHEYE', 8), 6: ('FULL_OPENCV', 12), 7: ('FOV', 5),
                  8: ('SIMPLE_RADIAL_FISHEYE', 4), 9: ('RADIAL_FISHEYE', 5), 10: ('THIN_PRISM_FISHEYE', 12)}


def quaternion_to_matrix(q) -> np.ndarray:
    """Rotation matrix (float64) of a COLMAP quaternion (w, x, y, z): normalised, then Eigen's
    Quaterniond::toRotationMatrix() formula, which is what pycolmap's `rotation.matrix()` evaluates
    (sfm.py:220)."""
    q = np.asarray(q, dtype=np.float64)
    w, x, y, z = q / np.linalg.norm(q)
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array([[1 - (tyy + tzz), txy - twz, txz + twy],
                     [txy + twz, 1 - (txx + tzz), tyz - twx],
                     [txz - twy, tyz + twx, 1 - (txx + tyy)]], dtype=np.float64)


def read_colmap_model(model_dir: Path) -> tuple[dict, dict]:
    """Minimal COLMAP reconstruction reader: returns (cameras, images) with
    cameras[id] = dict(model, width, height, params) and images[id] = dict(name, camera_id, qvec, tvec),
    keeping file order.  Binary files win over text, like pycolmap.Reconstruction."""
    model_dir = Path(model_dir)
    if (model_dir / 'cameras.bin').exists() and (model_dir / 'images.bin').exists():
        return _read_cameras_bin(model_dir / 'cameras.bin'), _read_images_bin(model_dir / 'images.bin')
    if (model_dir / 'cameras.txt').exists() and (model_dir / 'images.txt').exists():
        return _read_cameras_txt(model_dir / 'cameras.txt'), _read_images_txt(model_dir / 'images.txt')
    raise FileNotFoundError(f'no COLMAP model (cameras/images .bin or .txt) in {model_dir}')


def _read_cameras_txt(path: Path) -> dict:
    cameras = {}
    for line in path.read_text().splitlines():
        line = line.strip()
        if not line or line.startswith('#'):
            continue
        tok = line.split()
        cameras[int(tok[0])] = dict(model=tok[1], width=int(tok[2]), height=int(tok[3]),
                                    params=[float(x) for x in tok[4:]])
    return cameras


def _read_images_txt(path: Path) -> dict:
    images = {}
    expect_points = False
    for line in path.read_text().splitlines():
        if line.startswith('#'):
            continue
        if expect_points:          # second line of a record: 2D points, possibly empty
            expect_points = False
            continue
        tok = line.split()
        if not tok:
            continue
        images[int(tok[0])] = dict(qvec=[float(x) for x in tok[1:5]], tvec=[float(x) for x in tok[5:8]],
                                   camera_id=int(tok[8]), name=' '.join(tok[9:]))
        expect_points = True
    return images


def _read_cameras_bin(path: Path) -> dict:
    cameras = {}
    with open(path, 'rb') as f:
        (n,) = struct.unpack('<Q', f.read(8))
        for _ in range(n):
            cam_id, model_id, width, height = struct.unpack('<iiQQ', f.read(24))
            name, n_params = _CAMERA_MODELS[model_id]
            params = list(struct.unpack(f'<{n_params}d', f.read(8 * n_params)))
            cameras[cam_id] = dict(model=name, width=width, height=height, params=params)
    return cameras


def _read_images_bin(path: Path) -> dict:
    images = {}
    with open(path, 'rb') as f:
        (n,) = struct.unpack('<Q', f.read(8))
        for _ in range(n):
            image_id, *qt, camera_id = struct.unpack('<i7di', f.read(64))
            name = bytearray()
            while (ch := f.read(1)) != b'\x00':
                name += ch
            (n_pts,) = struct.unpack('<Q', f.read(8))
            f.seek(24 * n_pts, 1)
            images[image_id] = dict(qvec=qt[:4], tvec=qt[4:], camera_id=camera_id, name=name.decode())
    return images


class Pose:
    def __init__(self, R: Tensor, t: Tensor):
        """6 degree-of-freedom camera pose: rotation (3,3) and translation (3,1) (sfm.py:32-40)."""
        self.R = R
        self.t = t

    def inverse(self) -> Pose:
        return Pose(self.R.T, -self.R.T @ self.t)  # sfm.py:47, same expression => same bits

    def transform(self, P: Tensor) -> Tensor:
        return self.R.to(P.device) @ P + self.t.to(P.device)  # sfm.py:55

    def __repr__(self) -> str:
        return f'Pose(R={self.R!r}, t={self.t!r})'


class Camera:
    def __init__(self, camera_id: int, width: int, height: int, K: Tensor):
        self.id = camera_id
        self.width = width
        self.height = height
        self.K = K

    def __repr__(self) -> str:
        return f'Camera(id={self.id}, width={self.width}, height={self.height}, K={self.K!r})'


class Image:
    def __init__(self, image_id: int, rgb_path: Path, depth_map_path: Path, pose: Pose, camera: Camera):
        self.id = image_id
        self.name = str(rgb_path.name)  # basename, like sfm.py:84
        self.rgb_path = rgb_path
        self.depth_map_path = depth_map_path
        self.pose = pose
        self.camera = camera
        self.model: COLMAPModel | None = None  # set by COLMAPModel; owner of the device-resident scene
        self._geom: ViewGeom | None = None

    @property
    def geom(self) -> ViewGeom:
        if self._geom is None:
            self._geom = ViewGeom.from_pose(self.camera.K, self.pose.R, self.pose.t, self.camera.width,
                                            self.camera.height)
        return self._geom

    # reference-compatible float loaders (sfm.py:109-113); the hot path uses the raw integer ones below
    def get_rgb(self) -> Tensor:
        return loader.load_rgb(self.rgb_path, width=self.camera.width, height=self.camera.height)

    def get_depth_map(self) -> Tensor:
        return loader.load_depth_map(self.depth_map_path, width=self.camera.width, height=self.camera.height)

    def get_rgb_device_form(self) -> Tensor:
        return loader.load_rgb_device_form(self.rgb_path, width=self.camera.width, height=self.camera.height)

    def get_depth_u16(self) -> Tensor:
        return loader.load_depth_u16(self.depth_map_path, width=self.camera.width, height=self.camera.height)

    def match_images(self, image_list: list[Image], matches_file: loader.MatchesFile, min_cover: float = 0.000001,
                     num_workers: int = 0, device: str = 'cuda'):
        """Same contract as sfm.py:127-138, fused on the GPU: every target pixel is matched against every image of
        `image_list` with the two-way integer round-trip test, views covering <= min_cover of the target are
        dropped, and colour + range of the surviving source pixels are sampled (what the reference defers to
        loader.MatchesFile.prepare_matches / load_matches).  The result stays on the device, in `matches_file`.
        Kept views are consumed in name-sorted order, the order the reference's HDF5 groups iterate in
        (loader.py:63-66, 106)."""
        ordered = sorted(image_list, key=lambda im: im.name)
        keep = None
        if len(ordered) >= 128:
            # large surveys: decode the target first, drop the views its frustum cannot reach (conservative pre-test,
            # engine.DeviceScene.possibly_overlapping), and decode / upload only the others
            scene = self.model.scene(device, [self], num_workers=num_workers)
            keep = scene.possibly_overlapping(self.id, [im.id for im in ordered], [im.geom for im in ordered])
            if not keep.any():
                keep[0] = True
        needed = [im for i, im in enumerate(ordered) if keep is None or keep[i]]
        scene = self.model.scene(device, [self] + needed, num_workers=num_workers)
        store = gather(scene, self.id, [im.id for im in ordered], min_cover=min_cover, keep_src=True,
                       with_points=getattr(matches_file, 'with_points', False), keep_mask=keep)
        matches_file.set_store(store, [im.name for im in ordered])

    def __repr__(self) -> str:
        return f'SfMImage({self.name!r})'


class COLMAPModel:
    def __init__(self, model_dir: Path, image_dir: Path, depth_dir: Path, image_scale: float = 1.0):
        """Reads an undistorted COLMAP model; same attributes as the reference's (sfm.py:186-226)."""
        cameras, images = read_colmap_model(Path(model_dir))
        image_dir, depth_dir = Path(image_dir), Path(depth_dir)
        self.image_scale = image_scale

        self.cameras = {}
        for camera_id, cam in cameras.items():
            assert cam['model'] == 'PINHOLE', f'Camera {camera_id} is not using the PINHOLE model.'  # sfm.py:192
            width = int(cam['width'] * image_scale)
            height = int(cam['height'] * image_scale)
            scale_w = width / cam['width']
            scale_h = height / cam['height']
            fx, fy, u0, v0 = cam['params']
            fx, u0 = fx * scale_w, u0 * scale_w
            fy, v0 = fy * scale_h, v0 * scale_h
            self.cameras[camera_id] = Camera(camera_id=camera_id, width=width, height=height,
                                             K=torch.tensor([[fx, 0, u0], [0, fy, v0], [0, 0, 1]], dtype=torch.float32))

        self.images = {}
        for image_id, im in images.items():
            rgb_path = image_dir / im['name']
            depth_map_path = (depth_dir / im['name']).with_stem('depth_' + rgb_path.stem).with_suffix('.png')
            image = Image(
                image_id=image_id, rgb_path=rgb_path, depth_map_path=depth_map_path,
                pose=Pose(R=torch.tensor(quaternion_to_matrix(im['qvec']), dtype=torch.float32),
                          t=torch.tensor(im['tvec'], dtype=torch.float32).view(3, 1)).inverse(),  # cam->world
                camera=self.cameras[im['camera_id']])
            image.model = self
            self.images[image_id] = image

        self.imagename2id = {image.name: image.id for image in self.images.values()}
        self._scenes: dict = {}
        self._lru: dict = {}            # per device: view id -> tick of its last use
        self._tick = 0
        # device bytes the decoded views may occupy per GPU (None: 60 % of the device memory); beyond it the least
        # recently used views that the current target does not need are dropped and decoded again when needed
        self.scene_budget_bytes: int | None = None

    def __getitem__(self, image_name: str) -> Image:
        return self.images[self.imagename2id[image_name]]

    def __repr__(self) -> str:
        return f'COLMAPModel({len(self.images)} images)'

    def scene(self, device, images=None, num_workers: int = 0) -> DeviceScene:
        """The device-resident scene of this model on `device` (created on first use), with `images` (default:
        all) decoded and uploaded.  Decode happens once per image per device; a thread pool replaces the
        reference's DataLoader workers (loader.py:173-180)."""
        key = str(torch.device(device))
        if key not in self._scenes:
            self._scenes[key] = DeviceScene(device)
            self._lru[key] = {}
        scene, lru = self._scenes[key], self._lru[key]
        self._tick += 1
        todo, seen, wanted = [], set(), []
        for im in (self.images.values() if images is None else images):
            wanted.append(im.id)
            if im.id not in scene and im.id not in seen:
                todo.append(im)
                seen.add(im.id)
        for i in wanted:
            lru[i] = self._tick
        if todo:  # stay within the residency budget: drop the least recently used views this call does not ask for
            budget = self.scene_budget_bytes if self.scene_budget_bytes is not None else \
                int(0.6 * torch.cuda.get_device_properties(torch.device(device)).total_memory)
            per_view = {im.id: 5 * im.camera.width * im.camera.height for im in todo}
            resident = sum(scene.view_bytes(i) for i in scene.geom)
            excess = resident + sum(per_view.values()) - budget
            if excess > 0:
                keep = set(wanted)
                victims = []
                for i in sorted((i for i in scene.geom if i not in keep), key=lambda i: lru.get(i, 0)):
                    if excess <= 0:
                        break
                    excess -= scene.view_bytes(i)
                    victims.append(i)
                scene.remove_views(victims)
                for i in victims:
                    lru.pop(i, None)
        if todo:
            def decode(im):
                return im, im.get_depth_u16(), im.get_rgb_device_form()
            workers = num_workers if num_workers > 0 else min(8, os.cpu_count() or 1)  # cv2 decode releases the GIL
            if workers > 1 and len(todo) > 1:
                with ThreadPoolExecutor(max_workers=workers) as pool:
                    decoded = list(pool.map(decode, todo))
            else:
                decoded = [decode(im) for im in todo]
            for im, depth, rgb in decoded:
                scene.add_view(im.id, im.geom, depth, rgb)
        return scene

    def drop_scene(self, device=None):
        if device is None:
            self._scenes.clear()
            self._lru.clear()
        else:
            self._scenes.pop(str(torch.device(device)), None)
            self._lru.pop(str(torch.device(device)), None)
