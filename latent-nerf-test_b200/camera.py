"""Host-side camera math of the render path.

``generate_perspective_projection`` / ``generate_transformation_matrix`` stand where the
reference calls ``kal.render.camera.*`` (reference ``src/latent_paint/models/render.py:11,30``,
``src/latent_paint_mesh/models/render.py:18-19,54``); ``camera_from_view`` is the body of
``Renderer.get_camera_from_view`` (``render.py:19-31`` / ``:42-55``).

The latent_paint flavour of the reference computes its single camera on the CPU from Python
floats and only then moves the 12 numbers to the device; this module does the same with the
same torch ops, so that path is bit-identical to the reference glue.  Device-resident angle
tensors (latent_paint_mesh training loop) go through ``lp_cameras_from_views`` instead.
"""
from __future__ import annotations

import numpy as np
import torch


def generate_perspective_projection(fovyangle, ratio=1.0, dtype=torch.float):
    tanfov = np.tan(fovyangle / 2.0)
    return torch.tensor([[1.0 / (ratio * tanfov)], [1.0 / tanfov], [-1]], dtype=dtype)


def _unit(v):
    return v / torch.sqrt((v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]) + v[:, 2] * v[:, 2])[:, None]


def _cross(a, b):
    return torch.stack([a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1],
                        a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2],
                        a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]], dim=1)


def generate_transformation_matrix(camera_position, look_at, camera_up_direction):
    """(B,3) position, look-at point and up vector → (B,4,3) view matrix ``[R; t]`` with the
    camera axes as the columns of R and t = -pos·R (summed left to right)."""
    pos = camera_position.float()
    at = look_at.float().expand_as(pos)
    up = camera_up_direction.float().expand_as(pos)
    z = _unit(pos - at)
    x = _unit(_cross(up, z))
    y = _cross(z, x)
    rot = torch.stack([x, y, z], dim=2)
    t = -((pos[:, 0:1] * rot[:, 0, :] + pos[:, 1:2] * rot[:, 1, :]) + pos[:, 2:3] * rot[:, 2, :])
    return torch.cat([rot, t[:, None, :]], dim=1)


def camera_from_view(elev, azim, radius, look_at_height):
    """Spherical view parameters → (B,4,3).  ``elev`` is the polar angle from +y."""
    x = radius * torch.sin(elev) * torch.sin(azim)
    y = radius * torch.cos(elev)
    z = radius * torch.sin(elev) * torch.cos(azim)
    pos = torch.stack([torch.as_tensor(x, dtype=torch.float32).reshape(-1),
                       torch.as_tensor(y, dtype=torch.float32).reshape(-1),
                       torch.as_tensor(z, dtype=torch.float32).reshape(-1)], dim=1)
    at = torch.zeros_like(pos)
    at[:, 1] = torch.as_tensor(look_at_height, dtype=torch.float32).to(pos.device)
    up = torch.tensor([[0.0, 1.0, 0.0]], device=pos.device)
    return generate_transformation_matrix(pos, at, up)
