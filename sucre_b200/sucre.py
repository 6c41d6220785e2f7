"""SUCRe restoration with the reference's CLI and Python call surface (reference: sucre/sucre.py), running
the gather and the fit as sm_100a CUDA kernels.

    python -m sucre_b200.sucre --image-dir ... --depth-dir ... --model-dir ... --output-dir ... --image-name ...

Same flags, defaults and output files as the reference (sucre.py:264-307, 212-215).  Differences a user can see:
  * `--device` must be a CUDA device (there is no CPU path);
  * the matches file kept by `--keep-matches` is `<stem>.matches.npz` (+ an empty `<stem>.h5` marker), not HDF5;
  * with `--light-model` the ten light parameters are stepped on the host (one device->host read per iteration);
  * the per-iteration log lines are printed when the device loop returns, not while it runs.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch
from PIL import Image
from torch import Tensor
from tqdm import tqdm

from . import engine, light, loader, se3, sfm


class SUCRe:
    """Underwater image formation model I = J e^{-beta z} + B (1 - e^{-gamma z}) (sucre.py:35-82).

    Attributes B, beta, gamma are (3,1) tensors like the reference's Parameters (views of one 9-float device
    buffer); J is (H,W,3)."""

    def __init__(self, image: sfm.Image, light_model: bool = False, use_closed_form: bool = False):
        self.image = image
        self.light_model = light_model
        self.use_closed_form = use_closed_form
        J0 = None
        if not use_closed_form:
            depth, rgb = self._target_arrays()
            J0 = rgb.clone() if rgb.dtype == torch.float32 else rgb.to(torch.float32) / 255.0   # sucre.py:48
            J0[depth.view(torch.int16) == 0] = torch.nan                                       # sucre.py:49
        self.state = engine.FitState.initial('cpu' if J0 is None else J0.device, J0=J0)
        self._J_closed: Tensor | None = None
        self.history: Tensor | None = None
        if light_model:  # sucre.py:44-46; these ten scalars live on the host (see light.py)
            self.cam2light = torch.zeros(6)
            self.sigma = torch.eye(2)

    def _target_arrays(self) -> tuple[Tensor, Tensor]:
        """(u16 depth, colour in device form) of the target: taken from the model's device-resident scene when the
        image has already been decoded there (the usual case: the gather ran first), decoded from disk otherwise."""
        model = getattr(self.image, 'model', None)
        for scene in (model._scenes.values() if model is not None else ()):
            if self.image.id in scene and self.image.id in scene.rgb:
                return scene.depth[self.image.id], scene.rgb[self.image.id]
        return self.image.get_depth_u16(), self.image.get_rgb_device_form()

    # -- parameters -------------------------------------------------------------------------------------------
    @property
    def B(self) -> Tensor:
        return self.state.params[0:3].view(3, 1)

    @property
    def beta(self) -> Tensor:
        return self.state.params[3:6].view(3, 1)

    @property
    def gamma(self) -> Tensor:
        return self.state.params[6:9].view(3, 1)

    @property
    def J(self) -> Tensor | None:
        """Restored image (H,W,3): the Adam parameter in the default mode, the last update_J result otherwise."""
        return self._J_closed if self.use_closed_form else self.state.J

    @property
    def device(self) -> torch.device:
        return self.state.params.device

    def to(self, device) -> SUCRe:
        st = self.state
        st.params, st.moments = st.params.to(device), st.moments.to(device)
        if st.J is not None:
            st.J = st.J.to(device)
        if st.J_moments is not None:
            st.J_moments = st.J_moments.to(device)
        if self._J_closed is not None:
            self._J_closed = self._J_closed.to(device)
        return self

    def cpu(self) -> SUCRe:
        return self.to('cpu')

    def state_dict(self) -> dict:
        sd = {'B': self.B.detach().clone(), 'beta': self.beta.detach().clone(), 'gamma': self.gamma.detach().clone()}
        if self.light_model:
            sd['cam2light'] = self.cam2light.detach().clone()
            sd['sigma'] = self.sigma.detach().clone()
        if not self.use_closed_form:
            sd['J'] = self.J.detach().clone()
        return sd

    def state_dict_cpu(self) -> dict:
        """state_dict() with host tensors, leaving the model on its device (the plots are rendered there afterwards)."""
        return {k: v.cpu() for k, v in self.state_dict().items()}

    def load_state_dict(self, state_dict: dict, strict: bool = True):
        known = {'B': 0, 'beta': 3, 'gamma': 6}
        for key, value in state_dict.items():
            if key in known:
                self.state.params[known[key]:known[key] + 3] = torch.as_tensor(value, dtype=torch.float32).reshape(3).to(self.device)
            elif key == 'J' and not self.use_closed_form:
                self.state.J = torch.as_tensor(value, dtype=torch.float32).to(self.device).contiguous()
            elif key in ('cam2light', 'sigma') and self.light_model:
                setattr(self, key, torch.as_tensor(value, dtype=torch.float32).cpu().clone())
            elif strict:
                raise KeyError(f'unexpected key {key!r} in state_dict')

    # -- model ------------------------------------------------------------------------------------------------
    def compute_l_z(self, cP: Tensor) -> tuple[float | Tensor, Tensor]:
        """sucre.py:52-64 (torch ops; used by the plots, the fit evaluates this inside its kernels)."""
        z = cP.norm(dim=0)
        if not self.light_model:
            return 1.0, z
        R, t = se3.exp(self.cam2light)
        Sigma = self.sigma.T @ self.sigma
        lP = R.to(cP.device) @ cP + t.to(cP.device)
        lp = (lP[:2] / lP[2]).T.unsqueeze(dim=2)
        l = torch.exp(-torch.flatten(lp.transpose(1, 2) @ Sigma.inverse().to(cP.device) @ lp) / 2)
        return l, z + lP.norm(dim=0)

    def _light_params24(self) -> Tensor:
        return light.derive(self.B.cpu(), self.beta.cpu(), self.gamma.cpu(), self.cam2light, self.sigma).to(self.device)

    @torch.no_grad()
    def update_J(self, matches_data: loader.MatchesData, force_update: bool = False):
        """Closed-form J from the current B, beta, gamma (sucre.py:66-77) — one CUDA kernel over the store."""
        if self.light_model:
            if self.use_closed_form or force_update:
                J = engine.light_J(matches_data.store, self._light_params24())
                if self.use_closed_form:
                    self._J_closed = J
                else:
                    self.state.J = J
        elif self.use_closed_form:
            self._J_closed = engine.closed_form_J(matches_data.store, self.state.params, self.state.J)
        elif force_update:
            self.state.J = engine.closed_form_J(matches_data.store, self.state.params, None)

    @torch.no_grad()
    def forward(self, u: Tensor, v: Tensor, cP: Tensor) -> Tensor:
        """I_hat (3,n) at pixels (u,v) with camera-frame points cP (sucre.py:79-82).  Used for the reconstruction
        plot only; the fit evaluates the model inside its kernels."""
        l, z = self.compute_l_z(cP)
        return l * (self.J[v, u].T * torch.exp(-self.beta * z) + self.B * (1 - torch.exp(-self.gamma * z)))

    __call__ = forward

    def _target_depth_metres(self) -> Tensor:
        depth = self._target_arrays()[0].to(self.device)
        return (depth.view(torch.int16).to(torch.int32) & 0xffff).to(torch.float32) / 1000.0   # loader.py:167

    # -- outputs (sucre.py:84-121), host side, once per image ----------------------------------------------------
    @torch.no_grad()
    def plot_J(self) -> Image.Image:
        """Per-channel 1-99 percentile stretch of the valid pixels of J (sucre.py:84-94), evaluated on J's device
        (sort + numpy.percentile's default linear interpolation; torch.quantile refuses inputs above 2^24 elements, i.e.
        targets beyond 16.7 Mpixel): only the uint8 image crosses to the host."""
        J = self.J
        valid = ~torch.isnan(J).any(dim=2)
        Jv = J[valid]                                                        # (n, 3)
        lo, hi = _percentiles(Jv, 0.01), _percentiles(Jv, 0.99)
        Jv = torch.clamp(Jv, lo, hi)
        Jv = Jv - Jv.min(dim=0).values
        Jv = Jv / Jv.max(dim=0).values
        out = torch.zeros_like(J)
        out[valid] = Jv
        return Image.fromarray((out * 255).to(torch.uint8).cpu().numpy())

    @torch.no_grad()
    def plot_reconstruction(self) -> Image.Image:
        dev = self.device
        depth = self._target_depth_metres()
        v, u = torch.where(depth > 0)
        cp = torch.stack([u + 0.5, v + 0.5, torch.ones_like(u)])
        cP = self.image.geom.Kinv.to(dev) @ (depth[v, u] * cp)
        I_rec = torch.zeros((self.image.camera.height, self.image.camera.width, 3), device=dev)
        I_rec[v, u] = self(u=u, v=v, cP=cP).clip(0, 1).T
        return Image.fromarray(np.uint8(I_rec.cpu().numpy() * 255))

    @torch.no_grad()
    def plot_l(self) -> Image.Image:
        """Vignetting map of the light model, jet-coloured (sucre.py:96-104)."""
        dev = self.device
        depth = self._target_depth_metres()
        v, u = torch.where(depth > 0)
        cp = torch.stack([u + 0.5, v + 0.5, torch.ones_like(u)])
        cP = self.image.geom.Kinv.to(dev) @ (depth[v, u] * cp)
        l, _ = self.compute_l_z(cP)
        l_map = torch.zeros((self.image.camera.height, self.image.camera.width), device=dev)
        l_map[v, u] = l
        return Image.fromarray(np.uint8(_jet(l_map.cpu().numpy())[:, :, :3] * 255))

    def save_plots(self, save_dir: Path, iteration: int = None, writer: loader.AsyncWriter | None = None):
        """Same files as sucre.py:115-121.  The images are rendered now (on the device); with `writer` their PNG
        encoding happens on its threads."""
        save_path = (save_dir / self.image.name).with_suffix('.png')
        suffix = '' if iteration is None else f'_{iteration:04d}'
        jobs = [(self.plot_J(), save_path.with_stem(f'{save_path.stem}_rgb{suffix}')),
                (self.plot_reconstruction(), save_path.with_stem(f'{save_path.stem}_reconstruction{suffix}'))]
        if self.light_model:
            jobs.append((self.plot_l(), save_path.with_stem(f'{save_path.stem}_vignetting{suffix}')))
        for img, path in jobs:
            if writer is None:
                _save_png(img, path)
            else:
                writer.submit(_save_png, img, path)


def _save_png(img: Image.Image, path: Path):
    """Same pixels as `img.save(path)` (PNG is lossless); OpenCV's encoder at a low deflate level is 2.4x faster than
    PIL's default, and at survey scale (two PNGs per target, milliseconds of GPU work per target) encoding is what the
    host spends its time on."""
    import cv2
    a = np.asarray(img)
    if a.ndim == 3 and a.shape[2] == 3 and a.dtype == np.uint8:
        if not cv2.imwrite(str(path), a[:, :, ::-1], [cv2.IMWRITE_PNG_COMPRESSION, 1]):
            raise OSError(f'could not write {path}')
    else:
        img.save(path)


def _percentiles(x: Tensor, q: float) -> Tensor:
    """Per-column q-quantile of x (n, c) with numpy.percentile's default linear interpolation, any n."""
    xs = torch.sort(x, dim=0).values
    pos = q * (xs.shape[0] - 1)
    i0 = int(np.floor(pos))
    i1 = min(i0 + 1, xs.shape[0] - 1)
    return xs[i0] + (xs[i1] - xs[i0]) * (pos - i0)


def _jet(x: np.ndarray) -> np.ndarray:
    """RGBA of matplotlib's 'jet' colormap (256-entry table from its published anchor points); matplotlib itself is
    used when it is installed."""
    try:
        import matplotlib.pyplot as plt
        return plt.colormaps['jet'](x)
    except ImportError:
        pass
    anchors = {'r': [(0, 0), (0.35, 0), (0.66, 1), (0.89, 1), (1, 0.5)],
               'g': [(0, 0), (0.125, 0), (0.375, 1), (0.64, 1), (0.91, 0), (1, 0)],
               'b': [(0, 0.5), (0.11, 1), (0.34, 1), (0.65, 0), (1, 0)]}
    grid = np.linspace(0, 1, 256)
    lut = np.stack([np.interp(grid, *zip(*anchors[c])) for c in 'rgb'] + [np.ones(256)], axis=1)
    idx = np.clip((np.nan_to_num(np.asarray(x, dtype=np.float64)) * 256).astype(np.int64), 0, 255)
    return lut[idx]


def _fmt(x) -> str:
    """str() of a small float32 vector under np.printoptions(precision=4) — what the reference's log line shows
    (sucre.py:149-152) — without numpy's (slow) array printer for ordinary magnitudes."""
    x32 = np.asarray(x, dtype=np.float32)
    x = x32.astype(np.float64)
    ax = np.abs(x[x != 0])
    if not np.all(np.isfinite(x)) or (ax.size and (ax.max() >= 1e8 or ax.min() < 1e-4 or ax.max() / ax.min() > 1e3)):
        with np.printoptions(precision=4):
            return str(x32)
    ints, fracs = [], []
    for v in x32:  # numpy's own shortest-repr digit generation (Dragon4), element by element
        i, _, f = np.format_float_positional(v, precision=4, unique=True, fractional=True, trim='.').partition('.')
        ints.append(i)
        fracs.append(f)
    left, digits = max(map(len, ints)), max(map(len, fracs))
    return '[' + ' '.join(i.rjust(left) + '.' + f.ljust(digits) for i, f in zip(ints, fracs)) + ']'


def _log_history(history: np.ndarray, first_iteration: int, cost_column: int = 9):
    lines = [f'iter: {first_iteration + k:04d}, cost: {row[cost_column]:.4e}, B: {_fmt(row[0:3])}, '
             f'beta: {_fmt(row[3:6])}, gamma: {_fmt(row[6:9])}' for k, row in enumerate(history)]
    tqdm.write('\n'.join(lines))


def adam(
        sucre: SUCRe,
        matches_data: loader.MatchesData,
        lr: float = 0.05,
        num_iter: int = 200,
        batch_size: int = 1,
        save_dir: Path = None,
        save_interval: int = None,
        device: str = 'cuda'
) -> SUCRe:
    """Full-batch Adam on B, beta, gamma (and J in the default mode), sucre.py:124-157.  `batch_size` only
    grouped views for memory in the reference (gradients accumulate over all batches before the single
    optimizer.step()), so it has no effect here: every iteration streams the whole store once."""
    print(f'Solve least squares with Adam optimizer ({num_iter} iterations).')
    store = matches_data.store
    if sucre.light_model:
        return _adam_light(sucre, matches_data, lr, num_iter, save_dir, save_interval)
    # the reference builds a fresh torch.optim.Adam on every call (sucre.py:136): step counter and moments start at zero,
    # the parameters (and J) carry over
    sucre.state.reset_optimizer()
    # iterations after whose step the reference saves intermediate plots (sucre.py:153-154): it % save_interval == 0
    plot_its = list(range(0, num_iter, save_interval)) if save_dir is not None and save_interval else []
    histories = []
    done = 0
    for it in plot_its + [None]:
        end = num_iter if it is None else it + 1
        if it is not None and sucre.use_closed_form:
            # the plotted J is the one of the top of iteration `it` (pre-step parameters, sucre.py:141)
            if it > done:
                histories.append(engine.fit(store, sucre.state, it - done, lr))
            sucre.update_J(matches_data)
            histories.append(engine.fit(store, sucre.state, 1, lr))
        elif end > done:
            histories.append(engine.fit(store, sucre.state, end - done, lr))
        done = end
        if it is not None:
            sucre.save_plots(save_dir=save_dir, iteration=it)
    sucre.history = torch.cat(histories) if histories else None
    if sucre.history is not None:
        _log_history(sucre.history.cpu().numpy(), 0)
    sucre.update_J(matches_data=matches_data)  # sucre.py:156 (a no-op in the default mode, like the reference)
    return sucre


def _adam_light(sucre: SUCRe, matches_data: loader.MatchesData, lr: float, num_iter: int, save_dir, save_interval):
    """adam() with the light model: kernels for everything per-pixel, torch.optim.Adam on the 19 host scalars."""
    store = matches_data.store
    if not store.has_points:
        raise engine._lib.SucreError('the light model needs matches computed with camera-frame points; rerun with '
                                     '--force-compute-matches')
    dev = sucre.device
    names = ('B', 'beta', 'gamma', 'cam2light', 'sigma')
    params = {'B': sucre.B.cpu().clone(), 'beta': sucre.beta.cpu().clone(), 'gamma': sucre.gamma.cpu().clone(),
              'cam2light': sucre.cam2light.clone(), 'sigma': sucre.sigma.clone()}
    for p in params.values():
        p.requires_grad_(True)
    optimizer = torch.optim.Adam([params[k] for k in names], lr=lr)

    def sync_back():
        with torch.no_grad():
            sucre.state.params.copy_(torch.cat([params[k].flatten() for k in names[:3]]).to(dev))
            sucre.cam2light, sucre.sigma = params['cam2light'].detach().clone(), params['sigma'].detach().clone()

    plot_its = list(range(0, num_iter, save_interval)) if save_dir is not None and save_interval else []
    histories, done = [], 0
    J = None if sucre.use_closed_form else sucre.state.J
    for it in plot_its + [None]:
        end = num_iter if it is None else it + 1
        if end > done:
            h, J = light.fit(store, params, J, None if sucre.use_closed_form else sucre.state.J_moments, end - done, lr,
                             optimizer, first_step=done + 1)
            histories.append(h)
            done = end
        sync_back()
        if sucre.use_closed_form:
            sucre._J_closed = J
        if it is not None:
            sucre.save_plots(save_dir=save_dir, iteration=it)
    sucre.history = torch.cat(histories) if histories else None
    if sucre.history is not None:
        _log_history(sucre.history.numpy(), 0, cost_column=19)
    sucre.update_J(matches_data=matches_data)  # sucre.py:156
    return sucre


def restore_image(
        image: sfm.Image,
        colmap_model: sfm.COLMAPModel,
        output_dir: Path,
        light_model: bool = False,
        use_closed_form: bool = False,
        min_cover: float = 0.000001,
        image_list: list[sfm.Image] = None,
        lr: float = 0.05,
        num_iter: int = 200,
        batch_size: int = 1,
        save_interval: int = None,
        params_path: Path = None,
        force_compute_matches: bool = False,
        keep_matches: bool = False,
        num_workers: int = 0,
        device: str = 'cuda',
        *,
        writer: loader.AsyncWriter | None = None
):
    """Drop-in for the reference's restore_image (sucre.py:160-219): same arguments, same prints, same files.
    writer (keyword-only, optional): PNG encoding and the .pt dump are handed to its threads, so that the next target
    can start while this one's files are written; the caller closes it.  Without it every file exists on return."""
    print(f'Restore {image.name}.')
    matches_path = (output_dir / image.name).with_suffix('.h5')
    matches_file = loader.MatchesFile(matches_path, colmap_model=colmap_model, overwrite=force_compute_matches)

    if image_list is None:
        image_list = list(colmap_model.images.values())

    matches_file.with_points = light_model  # the light model needs cP, not only its norm (sucre.py:57)
    if force_compute_matches or not matches_file.exists():
        print(f'Compute {image.name} matches.')
        image.match_images(image_list=image_list, matches_file=matches_file, min_cover=min_cover,
                           num_workers=num_workers, device=device)
        print('Prepare matches for optimization.')
        matches_file.prepare_matches(num_workers=num_workers)

    print('Check matches integrity.')
    matches_file.check_integrity()

    print('Load matches.')
    matches_data = matches_file.load_matches(pin_memory=False, device=device)
    print(f'Total of {len(matches_data)} observations.')

    sucre = SUCRe(image=image, light_model=light_model, use_closed_form=use_closed_form).to(device)

    if params_path is not None:
        sucre.load_state_dict(torch.load(params_path), strict=False)

    adam(sucre=sucre, matches_data=matches_data, lr=lr, num_iter=num_iter, batch_size=batch_size,
         save_dir=output_dir, save_interval=save_interval, device=device)

    # the parameters first: a failure while rendering the plots must not lose the result of the fit
    J = sucre.J.detach().cpu()
    payload = {**sucre.state_dict_cpu(), 'J': J}
    if writer is None:
        torch.save(payload, (output_dir / image.name).with_suffix('.pt'))
    else:
        writer.submit(torch.save, payload, (output_dir / image.name).with_suffix('.pt'))
    sucre.save_plots(save_dir=output_dir, writer=writer)

    if keep_matches:
        matches_file.save()
    else:
        print(f'Erase {matches_path}.')
        matches_file.unlink()
    return None


from .cli import build_parser, main, parse_args  # noqa: E402,F401  (the CLI lives in cli.py; re-exported here)


if __name__ == '__main__':
    main()
