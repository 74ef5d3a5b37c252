"""Drop-in ``Renderer`` for ``src/latent_paint/models/render.py`` (reference lines 5-69).

Same constructor, method names, argument meaning, defaults, return shapes and dtypes as the
reference class; the work is done by the sm_100a kernels in ``csrc/lp_b200.cu`` instead of
kaolin + ATen.  Differences, all additive:
  * outputs are contiguous NCHW tensors (the reference returns permuted views);
  * ``renderer.last_buffers`` holds the visibility buffers of the last call when
    ``renderer.keep_buffers`` is set (face_idx int32, bary, depth, uv) — the reference
    returns no depth / face index, the tuple arity is unchanged;
  * all three interpolation modes of the reference's assert (nearest / bilinear / bicubic) are implemented; bicubic
    takes the split forward (footprint kernel + ``k_shade``), the other two the fused one.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, camera, functional


def decree_flags(r) -> int:
    """LP_FLAG_* bits of the decree switches set on a renderer object."""
    return (_lib.LP_FLAG_BBOX_HALF_OPEN if getattr(r, "bbox_half_open", False) else 0) | \
        (_lib.LP_FLAG_PLAIN_EPS if getattr(r, "plain_eps", False) else 0) | \
        (_lib.LP_FLAG_AFFINE_INTERP if getattr(r, "affine_interpolation", False) else 0) | \
        (_lib.LP_FLAG_SH_BAND1_XZY if getattr(r, "sh_band1_xzy", False) else 0)


class Renderer:
    def __init__(self, device, dim=(224, 224), interpolation_mode='nearest'):
        assert interpolation_mode in ['nearest', 'bilinear', 'bicubic'], f'no interpolation mode {interpolation_mode}'
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError("lp_b200 Renderer needs a CUDA device: there is no CPU path")
        _lib.lib()  # fail now, loudly, if the native library is missing
        self.interpolation_mode = interpolation_mode
        self.camera_projection = camera.generate_perspective_projection(np.pi / 3).to(self.device)
        self._proj = tuple(float(v) for v in camera.generate_perspective_projection(np.pi / 3).reshape(-1))
        self.dim = dim
        self.background = torch.ones(dim).to(self.device).float()
        self.reject_behind_camera = True   # BASELINE.md decree 3
        # the other open points of the kaolin restatement, False = the decree; the oracle has
        # the same switches (oracle/kaolin_shim.py), so a diff against real kaolin is a flag flip
        self.bbox_half_open = False        # bounding-box test x0 < xmax, y0 < ymax instead of <=
        self.plain_eps = False             # s + eps instead of s + copysign(eps, s)
        self.affine_interpolation = False  # screen-space instead of perspective-correct interpolation
        self.keep_buffers = False
        self.last_buffers = {}

    @staticmethod
    def get_camera_from_view(elev, azim, r=3.0, look_at_height=0.0):
        """(1,4,3) look-at matrix on the CPU, as in the reference (render.py:19-31)."""
        return camera.camera_from_view(elev, azim, r, look_at_height)

    def _flags(self, white_background=False):
        flags = _lib.LP_FLAG_MASK_IMAGE
        if white_background:
            flags |= _lib.LP_FLAG_WHITE_BACKGROUND
        if self.reject_behind_camera:
            flags |= _lib.LP_FLAG_REJECT_BEHIND
        return flags | decree_flags(self)

    def _config(self, verts, faces, elev, azim, radius, look_at_height, dims, white_background):
        cam = self.get_camera_from_view(torch.tensor(elev), torch.tensor(azim), r=radius,
                                        look_at_height=look_at_height).to(self.device)
        return functional.RenderConfig(
            verts=functional._f32(verts, self.device), faces=functional._faces_i32(faces, self.device, verts.shape[0]),
            cameras=cam.contiguous(), proj=self._proj, H=int(dims[1]), W=int(dims[0]),
            flags=self._flags(white_background), interp=self.interpolation_mode, want_buffers=self.keep_buffers)

    def depth_map(self, size=64, normalised=True):
        """``(B,1,size,size)`` depth input for depth-conditioned guidance (reference ``src/stable_diffusion_depth.py:302-319``)
        of the last render: inverse distance on the surface, 0 on the background, bicubic resize to ``size``, min-max
        normalised to [-1, 1] over the batch as ``train_step`` does.  Needs ``keep_buffers = True`` during the render."""
        if "depth" not in self.last_buffers or self.last_buffers["depth"] is None:
            raise RuntimeError("depth_map() needs renderer.keep_buffers = True before the render call")
        return functional.depth_for_guidance(self.last_buffers["depth"], size, normalised)

    def render_single_view(self, mesh, face_attributes, elev=0, azim=0, radius=2, look_at_height=0.0):
        """Per-face-vertex colours ``(1,F,3,Cf)`` → ``(image (1,Cf,H,W), mask (1,1,H,W))``
        (reference render.py:34-47); gradients flow into ``face_attributes``."""
        cfg = self._config(mesh.vertices, mesh.faces, elev, azim, radius, look_at_height, self.dim, False)
        image, mask, face_idx, bary, depth = functional.render_face_features(face_attributes.to(self.device), cfg)
        if self.keep_buffers:
            self.last_buffers = {"face_idx": face_idx, "bary": bary, "depth": depth, "camera": cfg.cameras}
        return image, mask

    def render_train_composed(self, verts, faces, uv_face_attr, texture_map, env_sphere, background_sphere_colors,
                              elev=0, azim=0, radius=2, look_at_height=0.0):
        """The training-time render of ``TexturedMeshModel.render_train`` (reference
        ``src/latent_paint/models/textured_mesh.py:195-212``) in one call: ``render_single_view_texture`` of the
        object, ``render_single_view`` of the environment sphere with the same camera, and
        ``pred_back * (1 - mask) + pred_features * mask`` fused into the second render.
        → ``(image, mask, background, foreground)``, gradients into ``texture_map`` and ``background_sphere_colors``."""
        tex_cfg = self._config(verts, faces, elev, azim, radius, look_at_height, self.dim, False)
        tex_cfg.face_uv = functional._f32(uv_face_attr, self.device).reshape(-1, 3, 2)
        feat_cfg = self._config(env_sphere.vertices, env_sphere.faces, elev, azim, radius, look_at_height, self.dim, False)
        return functional.render_composed(texture_map, background_sphere_colors.to(self.device), tex_cfg, feat_cfg)

    def render_single_view_texture(self, verts, faces, uv_face_attr, texture_map, elev=0, azim=0, radius=2,
                                   look_at_height=0.0, dims=None, white_background=False):
        """``(image (1,C,H,W), mask (1,1,H,W))`` of a UV-textured mesh (reference render.py:50-69);
        gradients flow into ``texture_map`` only."""
        dims = self.dim if dims is None else dims
        cfg = self._config(verts, faces, elev, azim, radius, look_at_height, dims, white_background)
        cfg.face_uv = functional._f32(uv_face_attr, self.device).reshape(-1, 3, 2)
        image, mask, uv, face_idx, bary, depth, _, _ = functional.render_texture(texture_map, cfg)
        if self.keep_buffers:
            self.last_buffers = {"face_idx": face_idx, "bary": bary, "depth": depth, "uv": uv, "camera": cfg.cameras}
        return image, mask
