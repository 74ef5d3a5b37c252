"""Drop-in ``Renderer`` for ``src/latent_paint_mesh/models/render.py`` (reference lines 5-279):
batched views, head / body projections, DIB-R style feature list (UV, ones, vertex normals),
bilinear texture fetch without masking, spherical-harmonic lighting.

Kept reference behaviour worth knowing: the ``look_at_height`` argument is
ignored in favour of the head/body table (render.py:182-190); the float ``mask`` is the
interpolated all-ones feature; the image is *not* multiplied by the mask, so uncovered pixels
sample — and back-propagate into — texel (row T-1, col 0); faces are dropped only when the z
of their unit normal is exactly 0 (the reference passes ``abs(n_z)`` to the back-face test,
render.py:237); vertex normals are averaged, not re-normalised.  The dead / broken methods
``render_single_view`` and ``render_single_view_texture_lighting`` of the reference file
 exist only as stubs that raise.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, camera, functional
from .render import decree_flags


class Renderer:
    def __init__(self, device, dim=(224, 224), interpolation_mode='nearest',
                 lights=torch.tensor([1.0, 0.0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0])):
        assert interpolation_mode in ['nearest', 'bilinear', 'bicubic'], f'no interpolation mode {interpolation_mode}'
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError("lp_b200 Renderer needs a CUDA device: there is no CPU path")
        _lib.lib()
        self.interpolation_mode = interpolation_mode
        head = camera.generate_perspective_projection(np.pi / 12)
        body = camera.generate_perspective_projection(np.pi / 4)
        self.camera_projection = [head.to(self.device), body.to(self.device)]
        self._proj = [tuple(float(v) for v in head.reshape(-1)), tuple(float(v) for v in body.reshape(-1))]
        self._look_at = [0.4, -0.3]
        self.look_at_height = torch.tensor([[0.4], [-0.3]]).to(self.device)
        self.dim = dim
        self.background = torch.ones(dim).to(self.device).float()
        self.lights = lights.unsqueeze(0).to(self.device)
        self._lights_flat = self.lights.reshape(-1).to(torch.float32).contiguous()
        self.reject_behind_camera = True
        # decree switches (see latent-nerf-test_b200/render.py): False = BASELINE.md's decree
        self.bbox_half_open = self.plain_eps = self.affine_interpolation = False
        self.sh_band1_xzy = False          # SH band-1 axis order (x, z, y) instead of the decreed (y, z, x)
        self.keep_buffers = False
        self.last_buffers = {}

    def get_camera_from_view(self, elev, azim, radius=3.0, look_at_height=0.0):
        """(B,4,3) look-at matrices on ``self.device`` (reference render.py:42-55).  Angle tensors
        on the GPU use the device kernel; CPU tensors use the same torch ops as the reference
        and are copied over (48 bytes per view)."""
        h = float(look_at_height.reshape(-1)[0]) if torch.is_tensor(look_at_height) and not look_at_height.is_cuda \
            else look_at_height
        if torch.is_tensor(elev) and elev.is_cuda:
            if torch.is_tensor(h):
                h = float(h.reshape(-1)[0])
            return functional.cameras_from_views(elev, azim, radius, h)
        elev, azim = torch.as_tensor(elev, dtype=torch.float32), torch.as_tensor(azim, dtype=torch.float32)
        if torch.is_tensor(radius):
            radius = radius.cpu()
        return camera.camera_from_view(elev, azim, radius, torch.as_tensor(h, dtype=torch.float32)).to(self.device)

    def compute_vertex_normals(self, faces, face_normals, num_vertices=None):
        """(B,F,3) unit face normals → (B,V,3) averaged vertex normals (reference render.py:57-105)."""
        import ctypes
        faces_i32 = functional._faces_i32(faces, self.device)
        V = int(faces.max()) + 1 if num_vertices is None else num_vertices
        off, vf = functional.vertex_face_csr(faces_i32, V)
        fn = functional._f32(face_normals, self.device)
        B, F = fn.shape[0], fn.shape[1]
        out = torch.empty((B, V, 3), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().lp_vertex_normals(functional._ptr(fn), functional._ptr(off), functional._ptr(vf),
                                                    B, V, F, functional._ptr(out), functional._stream(self.device)))
        return out

    def depth_map(self, size=64, normalised=True):
        """``(B,1,size,size)`` depth input for depth-conditioned guidance (reference ``src/stable_diffusion_depth.py:302-319``)
        of the last render: inverse distance on the surface, 0 on the background, bicubic resize to ``size``, min-max
        normalised to [-1, 1] over the batch as ``train_step`` does.  Needs ``keep_buffers = True`` during the render."""
        if "depth" not in self.last_buffers or self.last_buffers["depth"] is None:
            raise RuntimeError("depth_map() needs renderer.keep_buffers = True before the render call")
        return functional.depth_for_guidance(self.last_buffers["depth"], size, normalised)

    def render_single_view(self, *args, **kwargs):
        raise NotImplementedError("dead code in the reference (latent_paint_mesh/models/render.py:107-157 unpacks "
                                  "three tensors from kaolin's two-tuple rasterize and has no caller)")

    def render_single_view_texture_lighting(self, *args, **kwargs):
        raise NotImplementedError("broken in the reference (latent_paint_mesh/models/render.py:353 reads an undefined "
                                  "name; its only caller is commented out)")

    def render_single_view_texture(self, verts, faces, uv_face_attr, texture_map, elev=0, azim=0, radius=2,
                                   look_at_height=0.0, dims=None, white_background=False, disp=None, is_body=True):
        """→ ``(image (B,C,H,W), mask (B,1,H,W), normals (B,3,H,W), lighting (B,1,H,W))``
        (reference render.py:160-279); gradients flow into ``texture_map`` only."""
        dims = self.dim if dims is None else dims
        if disp is not None:
            verts = verts + disp
        P = 1 if is_body is True else 0
        cam = self.get_camera_from_view(elev, azim, radius, self._look_at[P])
        flags = _lib.LP_FLAG_CULL_NZ_ZERO
        if white_background:
            flags |= _lib.LP_FLAG_WHITE_BACKGROUND
        if self.reject_behind_camera:
            flags |= _lib.LP_FLAG_REJECT_BEHIND
        flags |= decree_flags(self)
        cfg = functional.RenderConfig(
            verts=functional._f32(verts, self.device), faces=functional._faces_i32(faces, self.device, verts.shape[0]),
            cameras=cam.contiguous(), proj=self._proj[P], H=int(dims[1]), W=int(dims[0]), flags=flags,
            interp='bilinear',                                           # hard-coded in the reference, render.py:243
            face_uv=functional._f32(uv_face_attr, self.device).reshape(-1, 3, 2),
            lights=self._lights_flat, want_buffers=self.keep_buffers)
        image, mask, uv, face_idx, bary, depth, normals, lighting = functional.render_texture(texture_map, cfg)
        if self.keep_buffers:
            self.last_buffers = {"face_idx": face_idx, "bary": bary, "depth": depth, "uv": uv, "camera": cam}
        return image, mask, normals, lighting
