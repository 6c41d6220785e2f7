"""TEST INFRASTRUCTURE ONLY — the reference's CPU algorithm restated op for op with PyTorch ATen calls.

Purpose: the CPU baseline of bench.py (`cpu_baseline` and `--impl reference`).  /root/reference does not exist on the
GPU box, so what is timed there is this port: the same whole-tensor ATen operations, in the same order, with the
same intermediates (stream compaction with torch.where, dense reverse index map, boolean-mask gathers, autograd +
torch.optim.Adam), only without the HDF5 spill file (matches stay in RAM).  It is NOT used as a parity checker on
the GPU (the C oracle is); tests/test_torch_port.py pins it bit-for-bit to the unmodified reference's golden
outputs on the CPU box.

Mapping to the reference (file:line in /root/reference/sucre/):
  backproject / to_pixels      sfm.py:90-107 (unproject_depth, unproject_depth_map, project_to_view) + sfm.py:42-55
  forward_matches, mutual      sfm.py:115-125, 154-159, 171-175
  gather_view, gather          sfm.py:127-138 + loader.py:78-87, 103-118
  FormationModel, run_adam     sucre.py:35-82, 124-157
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import Tensor


@dataclass
class PortView:
    K: Tensor       # (3,3) f32
    R: Tensor       # (3,3) cam->world
    t: Tensor       # (3,1)
    width: int
    height: int
    depth: Tensor   # (H,W) f32 metres  = f32(f64(u16) / 1000)   loader.py:167-170
    rgb: Tensor | None  # (H,W,3) f32    = f32(f64(u8) / 255)     loader.py:157-163


def make_view(K, R, t, width, height, depth_u16, rgb_u8=None) -> PortView:
    d = torch.tensor(depth_u16.numpy().astype('float64') / 1000, dtype=torch.float32)
    c = None if rgb_u8 is None else torch.tensor(rgb_u8.numpy().astype('float64') / 255, dtype=torch.float32)
    return PortView(K=K, R=R, t=t, width=int(width), height=int(height), depth=d, rgb=c)


def lift(view: PortView, u: Tensor, v: Tensor, d: Tensor) -> Tensor:
    pix = torch.stack([u + 0.5, v + 0.5, torch.ones_like(u)])
    return view.K.inverse() @ (d * pix)


def backproject(view: PortView):
    """Valid pixels (row-major, torch.where order) and their world points."""
    v, u = torch.where(view.depth > 0)
    cam = lift(view, u, v, view.depth[v, u])
    return u, v, view.R @ cam + view.t


def to_pixels(view: PortView, world: Tensor) -> Tensor:
    Ri = view.R.T
    ti = -view.R.T @ view.t
    cam = Ri @ world + ti
    img = view.K @ cam
    return (img[:2] / img[2]).long()


def forward_matches(dst: PortView, u1, v1, world):
    u2, v2 = to_pixels(dst, world)
    ok = (0 <= u2) & (u2 < dst.width) & (0 <= v2) & (v2 < dst.height)
    return u1[ok], v1[ok], u2[ok], v2[ok]


def mutual(target: PortView, source: PortView, pts_t, pts_s):
    """Two-way check through a dense reverse map of the source image."""
    a_u1, a_v1, a_u2, a_v2 = forward_matches(source, *pts_t)
    b_u1, b_v1, b_u2, b_v2 = forward_matches(target, *pts_s)   # b_*1 live in the source, b_*2 in the target
    rev = torch.full((source.height, source.width, 2), -1, dtype=b_u1.dtype)
    rev[b_v1, b_u1, 0] = b_v2
    rev[b_v1, b_u1, 1] = b_u2
    same = torch.all(rev[a_v2, a_u2] == torch.stack([a_v1, a_u1]).T, dim=1)
    return a_u1[same], a_v1[same], a_u2[same], a_v2[same]


def gather_view(target: PortView, source: PortView, pts_t, min_cover: float):
    """One (target, source) pair: matches, min_cover test, then the observation payload the reference builds in
    prepare_matches / load_matches.  Returns None when the view is dropped."""
    pts_s = backproject(source)
    u1, v1, u2, v2 = mutual(target, source, pts_t, pts_s)
    if not len(u1) / (target.width * target.height) > min_cover:
        return None
    d2 = source.depth[v2, u2]
    u1s, v1s, u2s, v2s = u1.short(), v1.short(), u2.short(), v2.short()   # the spill file stores int16
    I = source.rgb[v2s.long(), u2s.long()].T.contiguous()
    cP = lift(source, u2s, v2s, d2)
    return dict(u=u1s, v=v1s, u2=u2s, v2=v2s, d=d2, cP=cP, I=I)


def gather(target: PortView, sources: list[tuple[str, PortView]], min_cover: float = 1e-6):
    pts_t = backproject(target)
    kept = []
    for name, src in sources:
        obs = gather_view(target, src, pts_t, min_cover)
        if obs is not None:
            kept.append((name, obs))
    kept.sort(key=lambda kv: kv[0])  # HDF5 groups iterate sorted by name
    return kept


class FormationModel(torch.nn.Module):
    def __init__(self, height: int, width: int, closed_form: bool, J0: Tensor | None = None):
        super().__init__()
        self.hw = (height, width)
        self.closed_form = closed_form
        self.B = torch.nn.Parameter(torch.tensor([[0.1], [0.1], [0.1]]))
        self.beta = torch.nn.Parameter(torch.tensor([[0.1], [0.1], [0.1]]))
        self.gamma = torch.nn.Parameter(torch.tensor([[0.1], [0.1], [0.1]]))
        if not closed_form:
            self.J = torch.nn.Parameter(J0.clone())

    @torch.no_grad()
    def solve_J(self, views):
        num = torch.zeros((*self.hw, 3))
        den = torch.zeros((*self.hw, 3))
        for o in views:
            u, v = o['u'].long(), o['v'].long()
            z = o['cP'].norm(dim=0)
            att = 1.0 * torch.exp(-self.beta * z)
            back = 1.0 * self.B * (1 - torch.exp(-self.gamma * z))
            num[v, u] += ((o['I'] - back) * att).T
            den[v, u] += att.square().T
        self.J = num / den

    def forward(self, u, v, cP):
        z = cP.norm(dim=0)
        return 1.0 * (self.J[v, u].T * torch.exp(-self.beta * z) + self.B * (1 - torch.exp(-self.gamma * z)))


def adam_iteration(model: FormationModel, views: list[dict], opt, n_obs: int, batch_size: int = 5) -> float:
    """One iteration of adam() (sucre.py:138-148): zero_grad, update_J, view batches, step.  Returns the cost."""
    cost = 0.0
    opt.zero_grad()
    if model.closed_form:
        model.solve_J(views)
    for i in range(0, len(views), batch_size):
        chunk = views[i:i + batch_size]
        u = torch.hstack([o['u'] for o in chunk]).long()
        v = torch.hstack([o['v'] for o in chunk]).long()
        cP = torch.hstack([o['cP'] for o in chunk])
        I = torch.hstack([o['I'] for o in chunk])
        loss = torch.square(I - model(u, v, cP)).sum()
        (loss / n_obs / 3).backward()
        cost += loss.item()
    opt.step()
    return cost


def run_adam(model: FormationModel, views: list[dict], num_iter: int, lr: float = 0.05, batch_size: int = 5):
    """Returns (history (num_iter, 9), cost (num_iter,))."""
    n_obs = sum(o['u'].shape[0] for o in views)
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    history, costs = [], []
    for _ in range(num_iter):
        costs.append(adam_iteration(model, views, opt, n_obs, batch_size))
        history.append(torch.cat([model.B.detach().flatten(), model.beta.detach().flatten(),
                                  model.gamma.detach().flatten()]).clone())
    if model.closed_form:
        model.solve_J(views)
    return torch.stack(history), torch.tensor(costs, dtype=torch.float64)
