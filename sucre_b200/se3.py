"""SE(3) exponential used by the light model (reference: sucre/se3.py:22-27)."""
from __future__ import annotations

import torch
from torch import Tensor


def exp(pose: Tensor) -> tuple[Tensor, Tensor]:
    """(R (3,3), t (3,1)) = matrix exponential of the twist `pose` = (w1, w2, w3, p1, p2, p3)."""
    w, p = pose[:3], pose[3:]
    zero = torch.zeros((), dtype=pose.dtype, device=pose.device)
    twist = torch.stack([torch.stack([zero, -w[2], w[1], p[0]]),
                         torch.stack([w[2], zero, -w[0], p[1]]),
                         torch.stack([-w[1], w[0], zero, p[2]]),
                         torch.stack([zero, zero, zero, zero])])
    T = torch.matrix_exp(twist)
    return T[:3, :3], T[:3, 3:4]
