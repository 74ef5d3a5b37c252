"""TEST INFRASTRUCTURE — travelling mirror of the reference's two ``Renderer`` classes on top
of ``oracle/kaolin_shim.py`` (CPU, torch).  The GPU box has no /root/reference, so the GPU
parity tests, ``smoke()`` and the reference arm of ``bench.py`` use this mirror; here, in the
build container, ``tests/test_oracle.py`` pins it against the reference's real files
(``oracle/reference_glue.py``) and the frozen ``tests/golden/*.npz``.

It follows the reference step by step and cites it, but is written against a ``device``
parameter and returns the intermediate buffers the parity tests compare (face_idx, uv).
"""
from __future__ import annotations

import numpy as np
import torch

from . import kaolin_shim as _oracle_kal


def _look_at_camera(elev, azim, radius, look_at_height, kal=_oracle_kal):
    """Camera position on the view sphere and the kaolin look-at matrix.
    reference latent_paint/models/render.py:19-31 and latent_paint_mesh/models/render.py:42-55."""
    x = radius * torch.sin(elev) * torch.sin(azim)
    y = radius * torch.cos(elev)
    z = radius * torch.sin(elev) * torch.cos(azim)
    pos = torch.stack([torch.as_tensor(x, dtype=torch.float32).reshape(-1),
                       torch.as_tensor(y, dtype=torch.float32).reshape(-1),
                       torch.as_tensor(z, dtype=torch.float32).reshape(-1)], dim=1)
    at = torch.zeros_like(pos)
    at[:, 1] = torch.as_tensor(look_at_height, dtype=torch.float32)
    up = torch.tensor([[0.0, 1.0, 0.0]])
    return kal.generate_transformation_matrix(pos, at, up)


def _ns(kal):
    """Accept either this package's flat shim module or a ``kaolin``-shaped module tree (the product's
    ``kaolin_compat`` — the same glue then runs on the CUDA kernels)."""
    if kal is None:
        return _oracle_kal
    if hasattr(kal, "render"):
        import types
        ns = types.SimpleNamespace()
        for name in ("generate_perspective_projection", "generate_transformation_matrix"):
            setattr(ns, name, getattr(kal.render.camera, name))
        for name in ("prepare_vertices", "rasterize", "dibr_rasterization", "texture_mapping", "spherical_harmonic_lighting"):
            setattr(ns, name, getattr(kal.render.mesh, name))
        ns.index_vertices_by_faces = kal.ops.mesh.index_vertices_by_faces
        return ns
    return kal


class LatentPaintRendererRef:
    """Mirror of reference ``src/latent_paint/models/render.py`` (single view, fov π/3)."""

    def __init__(self, dim=(224, 224), interpolation_mode="nearest", kal=None, device="cpu"):
        assert interpolation_mode in ["nearest", "bilinear", "bicubic"]        # render.py:9
        self.kal, self.device = _ns(kal), device
        kal = self.kal
        self.camera_projection = kal.generate_perspective_projection(np.pi / 3).to(device)  # render.py:11
        self.interpolation_mode = interpolation_mode
        self.dim = dim
        self.last = {}

    @staticmethod
    def get_camera_from_view(elev, azim, r=3.0, look_at_height=0.0):            # render.py:19-31
        return _look_at_camera(elev, azim, r, look_at_height)

    def render_single_view(self, vertices, faces, face_attributes, elev=0, azim=0, radius=2, look_at_height=0.0):
        """render.py:34-47 — per-face-vertex colours interpolated by the rasterizer."""
        dims, kal = self.dim, self.kal
        M = self.get_camera_from_view(torch.tensor(elev), torch.tensor(azim), r=radius, look_at_height=look_at_height).to(self.device)
        fvc, fvi, _ = kal.prepare_vertices(vertices, faces, self.camera_projection, camera_transform=M)
        feats, face_idx = kal.rasterize(dims[1], dims[0], fvc[:, :, :, -1], fvi, face_attributes)
        mask = (face_idx > -1).float()[..., None]
        self.last = {"face_idx": face_idx, "camera": M}
        return feats.permute(0, 3, 1, 2), mask.permute(0, 3, 1, 2)

    def render_single_view_texture(self, verts, faces, uv_face_attr, texture_map, elev=0, azim=0, radius=2,
                                   look_at_height=0.0, dims=None, white_background=False):
        """render.py:50-69."""
        dims, kal = (self.dim if dims is None else dims), self.kal
        M = self.get_camera_from_view(torch.tensor(elev), torch.tensor(azim), r=radius, look_at_height=look_at_height).to(self.device)
        fvc, fvi, _ = kal.prepare_vertices(verts, faces, self.camera_projection, camera_transform=M)
        uv, face_idx = kal.rasterize(dims[1], dims[0], fvc[:, :, :, -1], fvi, uv_face_attr)
        uv = uv.detach()                                                          # render.py:61
        mask = (face_idx > -1).float()[..., None]
        image = kal.texture_mapping(uv, texture_map, mode=self.interpolation_mode)
        image = image * mask
        if white_background:
            image = image + 1 * (1 - mask)
        self.last = {"face_idx": face_idx, "uv": uv, "camera": M}
        return image.permute(0, 3, 1, 2), mask.permute(0, 3, 1, 2)


class LatentPaintMeshRendererRef:
    """Mirror of reference ``src/latent_paint_mesh/models/render.py`` (batched views, head/body
    cameras, DIB-R feature list, SH lighting)."""

    def __init__(self, dim=(224, 224), interpolation_mode="nearest",
                 lights=torch.tensor([1.0, 0.0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0]), kal=None, device="cpu"):
        assert interpolation_mode in ["nearest", "bilinear", "bicubic"]        # render.py:16
        self.kal, self.device = _ns(kal), device
        kal = self.kal
        self.camera_projection = [kal.generate_perspective_projection(np.pi / 12).to(device),   # head, render.py:18
                                  kal.generate_perspective_projection(np.pi / 4).to(device)]    # body, render.py:19
        self.look_at_height = torch.tensor([[0.4], [-0.3]])                     # render.py:29-32
        self.interpolation_mode = interpolation_mode
        self.dim = dim
        self.lights = lights.unsqueeze(0).to(device)
        self.last = {}

    def get_camera_from_view(self, elev, azim, radius=3.0, look_at_height=0.0):  # render.py:42-55
        return _look_at_camera(elev, azim, radius, look_at_height, self.kal).to(self.device)

    @staticmethod
    def compute_vertex_normals(faces, face_normals, num_vertices=None):
        """render.py:57-105 — per corner k, scatter-add the (B,F,3) unit face normals onto the
        vertices, divide by the incidence count; NOT re-normalised."""
        V = int(faces.max()) + 1 if num_vertices is None else num_vertices
        B, F = face_normals.shape[0], faces.shape[0]
        dev = face_normals.device
        vn = torch.zeros((B, V, 3), dtype=face_normals.dtype, device=dev)
        cnt = torch.zeros((B, V), dtype=face_normals.dtype, device=dev)
        ones = torch.ones((B, F), dtype=face_normals.dtype, device=dev)
        for k in range(faces.shape[1]):
            vn.scatter_add_(1, faces[None, :, k:k + 1].repeat(B, 1, 3), face_normals)
            cnt.scatter_add_(1, faces[None, :, k].repeat(B, 1), ones)
        return vn / cnt.clip(min=1).unsqueeze(-1)

    def render_single_view_texture(self, verts, faces, uv_face_attr, texture_map, elev=0, azim=0, radius=2,
                                   look_at_height=0.0, dims=None, white_background=False, disp=None, is_body=True):
        """render.py:160-279.  ``look_at_height`` is ignored exactly as in the reference."""
        dims, kal = (self.dim if dims is None else dims), self.kal
        if disp is not None:
            verts = verts + disp
        P = 1 if is_body is True else 0
        M = self.get_camera_from_view(elev, azim, radius, self.look_at_height[P])
        B = M.shape[0]
        fvc, fvi, fn = kal.prepare_vertices(verts, faces, self.camera_projection[P], camera_transform=M)
        vn = self.compute_vertex_normals(faces, fn)
        vfn = kal.index_vertices_by_faces(vn, faces)
        feats = [uv_face_attr.repeat(B, 1, 1, 1), torch.ones((B, faces.shape[0], 3, 1), device=self.device), vfn]
        (uv, mask, normals), _soft, face_idx = kal.dibr_rasterization(
            dims[1], dims[0], fvc[:, :, :, -1], fvi, feats, abs(fn[:, :, -1]), rast_backend="cuda")
        image = kal.texture_mapping(uv, texture_map.repeat(B, 1, 1, 1), mode="bilinear")   # render.py:243
        lighting = kal.spherical_harmonic_lighting(normals, self.lights).unsqueeze(0)
        lighting = lighting.clamp(1e-8, 1).permute(1, 0, 2, 3)
        if white_background:
            image = image + 1 * (1 - mask)
        self.last = {"face_idx": face_idx, "uv": uv, "camera": M}
        return image.permute(0, 3, 1, 2), mask.permute(0, 3, 1, 2), normals.permute(0, 3, 1, 2), lighting


def render_train_ref(ref: "LatentPaintRendererRef", mesh_vertices, mesh_faces, face_attributes, texture_img,
                     env_vertices, env_faces, background_sphere_colors, theta, phi, radius, dy=0.25, latent_mode=True):
    """Mirror of ``TexturedMeshModel.render_train`` (reference ``src/latent_paint/models/textured_mesh.py:187-220``)
    over a renderer with the reference's ``Renderer`` interface (this module's mirror or the real class run by
    ``oracle/reference_glue.py``): object render, environment-sphere render, ``mask.detach()``, the composition and
    the bicubic resize to the 64 x 64 latent grid.  ``env_vertices`` / ``env_faces`` stand for ``self.env_sphere``."""
    import torch.nn.functional as F
    pred_features, mask = ref.render_single_view_texture(mesh_vertices, mesh_faces, face_attributes, texture_img,
                                                         elev=theta, azim=phi, radius=radius, look_at_height=dy)   # :195-202
    if isinstance(ref, LatentPaintRendererRef):
        pred_back, _ = ref.render_single_view(env_vertices, env_faces, background_sphere_colors, elev=theta, azim=phi,
                                              radius=radius, look_at_height=dy)                                  # :204-209
    else:                                      # the real reference Renderer takes a mesh object (render.py:34)
        import types
        pred_back, _ = ref.render_single_view(types.SimpleNamespace(vertices=env_vertices, faces=env_faces),
                                              background_sphere_colors, elev=theta, azim=phi, radius=radius, look_at_height=dy)
    mask = mask.detach()                                                                                         # :211
    pred_map = pred_back * (1 - mask) + pred_features * mask                                                     # :212
    if latent_mode and mask.shape[-1] != 64:                                                                      # :214-218
        mask = F.interpolate(mask, (64, 64), mode='bicubic')
        pred_back = F.interpolate(pred_back, (64, 64), mode='bicubic')
        pred_features = F.interpolate(pred_features, (64, 64), mode='bicubic')
        pred_map = F.interpolate(pred_map, (64, 64), mode='bicubic')
    return {'image': pred_map, 'mask': mask, 'background': pred_back, 'foreground': pred_features}
