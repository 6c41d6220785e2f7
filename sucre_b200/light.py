"""Host side of the light model (--light-model; reference: sucre/sucre.py:44-46, 54-61, 124-157).

The ten light parameters (cam2light twist, 2x2 sigma) enter the kernels only through R, t = se3.exp(cam2light) and
Sigma^-1 = (sigma^T sigma)^-1.  Those are evaluated here with the reference's own torch expressions; the kernels
(csrc/light.cu) return the sums that carry dL/dR, dL/dt and dL/dSigma^-1, and the chain rule back through
matrix_exp and the 2x2 inverse is one tiny autograd call per iteration.  The optimiser is torch.optim.Adam itself,
on the 19 host-resident scalars; per-pixel work (closed-form J, residual sums, J's own Adam step) stays on the GPU.
One device->host read of 25 doubles per iteration: this optional mode is correctness-first.
"""
from __future__ import annotations

import torch

from . import engine, se3


def derive(B, beta, gamma, cam2light, sigma) -> torch.Tensor:
    """The 24 floats the kernels take: B, beta, gamma, R (row-major), t, (Sigma^-1)_00, _01, _11."""
    R, t = se3.exp(cam2light)
    Sinv = (sigma.T @ sigma).inverse()
    return torch.cat([B.flatten(), beta.flatten(), gamma.flatten(), R.flatten(), t.flatten(),
                      torch.stack([Sinv[0, 0], Sinv[0, 1], Sinv[1, 1]])]).to(torch.float32)


def fit(store: engine.ObservationStore, params: dict, J: torch.Tensor | None, J_moments: torch.Tensor | None,
        num_iter: int, lr: float, optimizer: torch.optim.Optimizer, first_step: int = 1):
    """num_iter Adam iterations.  params: dict of CPU leaf tensors B, beta, gamma (3,1), cam2light (6,), sigma (2,2)
    registered in `optimizer`.  J_moments is None in closed-form mode (J is then recomputed every iteration from the
    pre-step parameters, sucre.py:141) and J's Adam state otherwise.  Returns (history (num_iter, 20) = the 19
    parameters after each step + cost, J of the last evaluated iteration)."""
    if store.n_obs == 0:
        raise engine._lib.SucreError('fit: the observation store is empty')
    dev = store.cells.device
    closed_form = J_moments is None
    sums = torch.zeros(25, dtype=torch.float64, device=dev)
    history = torch.zeros((num_iter, 20), dtype=torch.float32)
    sc = 2.0 / (3.0 * store.n_obs)
    names = ('B', 'beta', 'gamma', 'cam2light', 'sigma')
    for it in range(num_iter):
        # R, t, Sigma^-1 are evaluated once per iteration, with the graph the chain rule below walks back through
        xi = params['cam2light'].detach().clone().requires_grad_(True)
        sg = params['sigma'].detach().clone().requires_grad_(True)
        R, t = se3.exp(xi)
        Sinv = (sg.T @ sg).inverse()
        with torch.no_grad():
            p24 = torch.cat([params['B'].flatten(), params['beta'].flatten(), params['gamma'].flatten(), R.flatten(), t.flatten(),
                             torch.stack([Sinv[0, 0], Sinv[0, 1], Sinv[1, 1]])]).to(torch.float32).to(dev)
        if closed_form:
            J = engine.light_J(store, p24)
        engine.light_sums(store, p24, J, sums, J_moments, n_obs=store.n_obs, step=first_step + it, lr=lr)
        s = sums.cpu()  # the one host sync of the iteration
        optimizer.zero_grad()
        params['B'].grad = (-sc * s[0:3]).to(torch.float32).view(3, 1)
        params['beta'].grad = (sc * s[3:6]).to(torch.float32).view(3, 1)
        params['gamma'].grad = (-sc * s[6:9]).to(torch.float32).view(3, 1)
        # chain rule through se3.exp and the inverse: d/d(cam2light, sigma) of <dL/dR, R> + <dL/dt, t> + <dL/dS, S>
        dR = (-sc * s[13:22]).to(torch.float32).view(3, 3)
        dt = (-sc * s[22:25]).to(torch.float32).view(3, 1)
        dS = (0.5 * sc * torch.stack([torch.stack([s[10], s[11]]), torch.stack([s[11], s[12]])])).to(torch.float32)
        ((dR * R).sum() + (dt * t).sum() + (dS * Sinv).sum()).backward()
        params['cam2light'].grad = xi.grad
        params['sigma'].grad = sg.grad
        optimizer.step()
        with torch.no_grad():
            history[it, :19] = torch.cat([params[k].flatten() for k in names])
            history[it, 19] = float(s[9])
    return history, J
