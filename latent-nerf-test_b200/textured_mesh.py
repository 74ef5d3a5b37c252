"""Model-level render of Latent-Paint (SURVEY.md §8 f rank 1): what ``TexturedMeshModel.render_train`` does around
the renderer (reference ``src/latent_paint/models/textured_mesh.py:187-220``), over the fused composition.

Only the render call chain is mirrored; the model class itself (parameters, checkpoints, export) is out of scope.
"""
from __future__ import annotations

import torch

from . import functional


def render_train(renderer, mesh, face_attributes, texture_img, env_sphere, background_sphere_colors, theta, phi, radius,
                 dy: float = 0.25, latent_mode: bool = True, linear_rgb_estimator=None):
    """→ ``{'image', 'mask', 'background', 'foreground'}`` exactly as the reference returns them (:220).

    ``renderer`` is a ``LatentPaintRenderer``; ``mesh`` / ``env_sphere`` carry ``.vertices`` / ``.faces``;
    ``face_attributes`` are the object's per-face UVs ``(1,F,3,2)``; ``background_sphere_colors`` is the learnable
    ``(1,F_env,3,4)`` tensor.  In RGB fine-tuning mode (``latent_mode=False``) the sphere colours go through the
    4→3 linear estimator first (:191-193)."""
    if not latent_mode:
        if linear_rgb_estimator is None:
            raise ValueError("latent_mode=False needs the (4,3) linear_rgb_estimator")
        background_sphere_colors = background_sphere_colors @ linear_rgb_estimator
    pred_map, mask, pred_back, pred_features = renderer.render_train_composed(
        mesh.vertices, mesh.faces, face_attributes, texture_img, env_sphere, background_sphere_colors,
        elev=theta, azim=phi, radius=radius, look_at_height=dy)
    mask = mask.detach()
    if latent_mode and mask.shape[-1] != 64:
        # the reference resizes to the 64 x 64 latent grid with four bicubic F.interpolate calls (:214-218); here one
        # launch (lp_resize_bicubic: ATen's upsample_bicubic2d arithmetic) over all four tensors, forward and backward
        mask, pred_back, pred_features, pred_map = functional.resize_bicubic([mask, pred_back, pred_features, pred_map], (64, 64))
    return {'image': pred_map, 'mask': mask, 'background': pred_back, 'foreground': pred_features}
