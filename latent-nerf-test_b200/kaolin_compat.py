"""``kaolin``-namespaced operator API on top of the sm_100a kernels (SURVEY.md §8 b, second row).

Exposes exactly the kaolin entry points the reference imports, so that the reference's own
``src/latent_paint*/models/render.py`` (``import kaolin as kal``) runs unmodified with no kaolin
installed: call ``install()`` before importing them.  The heavy ops — ``rasterize`` /
``dibr_rasterization`` (+ their backward into the face features) and ``texture_mapping`` (+ its
backward into the texture maps) — are the CUDA kernels of ``csrc/lp_b200.cu``; the small per-vertex
ops are plain torch on the caller's device, written in the oracle's fixed expression order.

Reference call sites: see the table in ``oracle/kaolin_shim.py`` (the CPU oracle of the same API).
"""
from __future__ import annotations

import sys
import types

import torch

from . import _lib, camera, functional, meshio

REJECT_BEHIND_CAMERA = True      # BASELINE.md decree 3
# the other open points of the restatement as switches, mirrored in oracle/kaolin_shim.py (False = the decree)
BBOX_HALF_OPEN = False
PLAIN_EPS = False
AFFINE_INTERP = False


# ---------------------------------------------------------------- render.camera
generate_perspective_projection = camera.generate_perspective_projection
generate_transformation_matrix = camera.generate_transformation_matrix


# ---------------------------------------------------------------- ops.mesh
def index_vertices_by_faces(vertices_features, faces):
    B, _, K = vertices_features.shape
    F = faces.shape[0]
    idx = faces.reshape(1, F * 3, 1).expand(B, F * 3, K)
    return torch.gather(vertices_features, 1, idx).reshape(B, F, 3, K)


def uniform_laplacian(num_vertices, faces):
    """``kal.ops.mesh.uniform_laplacian``: dense (V,V), as the reference calls it (textured_mesh.py:62).  Use
    :func:`uniform_laplacian_sparse` for anything beyond a few thousand vertices."""
    adj = torch.zeros((num_vertices, num_vertices), dtype=torch.float32, device=faces.device)
    for a, b in ((0, 1), (1, 2), (2, 0)):
        adj[faces[:, a], faces[:, b]] = 1
        adj[faces[:, b], faces[:, a]] = 1
    deg = adj.sum(dim=1, keepdim=True).clamp(min=1)
    return adj / deg - torch.eye(num_vertices, device=faces.device)


def uniform_laplacian_sparse(num_vertices, faces):
    """The same operator as a sparse CSR tensor: 1/deg(i) on the neighbours of vertex i, -1 on the diagonal — 7 entries
    per vertex on a closed triangle mesh instead of V (the dense form of the reference, latent_paint_mesh
    textured_mesh.py:60-71, is 1.7 TB at config 4's 655 362 vertices).  ``laplacian_coordinates`` / ``lap_loss`` below are
    the two places the reference uses it (``L.mm(vertices)``, :65 and :314-317)."""
    f = faces.long()
    src = torch.cat([f[:, 0], f[:, 1], f[:, 1], f[:, 2], f[:, 2], f[:, 0]])
    dst = torch.cat([f[:, 1], f[:, 0], f[:, 2], f[:, 1], f[:, 0], f[:, 2]])
    key = torch.unique(src * num_vertices + dst)                       # each undirected edge once per direction
    row, col = key // num_vertices, key % num_vertices
    deg = torch.bincount(row, minlength=num_vertices).clamp(min=1).to(torch.float32)
    diag = torch.arange(num_vertices, device=faces.device)
    rows = torch.cat([row, diag]); cols = torch.cat([col, diag])
    vals = torch.cat([1.0 / deg[row], -torch.ones(num_vertices, device=faces.device)])
    return torch.sparse_coo_tensor(torch.stack([rows, cols]), vals, (num_vertices, num_vertices)).coalesce().to_sparse_csr()


def laplacian_coordinates(L, vertices):
    """``L.mm(vertices)`` for the dense or the sparse operator."""
    return torch.sparse.mm(L, vertices) if L.layout != torch.strided else L.mm(vertices)


def lap_loss(L, vertices, init_lap):
    """``torch.mean(torch.sum((L.mm(v) - init_lap) ** 2))`` (reference latent_paint_mesh textured_mesh.py:314-317)."""
    return torch.mean(torch.sum((laplacian_coordinates(L, vertices) - init_lap) ** 2))


# ---------------------------------------------------------------- render.mesh
def prepare_vertices(vertices, faces, camera_proj, camera_rot=None, camera_trans=None, camera_transform=None):
    """→ face_vertices_camera (B,F,3,3), face_vertices_image (B,F,3,2), unit face normals (B,F,3)."""
    if camera_transform is None:
        raise NotImplementedError("only the camera_transform form is used by the reference")
    v = vertices.float()
    if v.dim() == 2:
        v = v[None]
    M = camera_transform.float().to(v.device)
    vx, vy, vz = v[..., 0:1], v[..., 1:2], v[..., 2:3]
    cam = ((vx * M[:, None, 0, :] + vy * M[:, None, 1, :]) + vz * M[:, None, 2, :]) + M[:, None, 3, :]
    pp = cam * camera_proj.float().to(v.device).reshape(1, 1, 3)
    img = pp[..., :2] / pp[..., 2:3]
    fvc = index_vertices_by_faces(cam, faces)
    fvi = index_vertices_by_faces(img, faces)
    e0, e1 = fvc[:, :, 1] - fvc[:, :, 0], fvc[:, :, 2] - fvc[:, :, 0]
    n = torch.stack([e0[..., 1] * e1[..., 2] - e0[..., 2] * e1[..., 1],
                     e0[..., 2] * e1[..., 0] - e0[..., 0] * e1[..., 2],
                     e0[..., 0] * e1[..., 1] - e0[..., 1] * e1[..., 0]], dim=-1)
    ln = torch.sqrt((n[..., 0] * n[..., 0] + n[..., 1] * n[..., 1]) + n[..., 2] * n[..., 2])
    return fvc, fvi, n / (ln[..., None] + 1e-10)


def rasterize(height, width, face_vertices_z, face_vertices_image, face_features, valid_faces=None,
              multiplier=None, eps=None, backend="cuda"):
    """→ (interpolated_features (B,H,W,D) or a tuple of them, face_idx (B,H,W) int64).  Differentiable in
    the face features (kaolin's rasterize backward); not in the vertices (the reference never needs it)."""
    device = face_vertices_z.device
    functional._require_cuda(face_vertices_z, "face_vertices_z")
    is_list = isinstance(face_features, (list, tuple))
    feats = list(face_features) if is_list else [face_features]
    B = face_vertices_z.shape[0]
    feats = [f.to(device).float().expand(B, -1, -1, -1) if f.shape[0] == 1 and B > 1 else f.to(device).float() for f in feats]
    dims = [f.shape[-1] for f in feats]
    ff = torch.cat(feats, dim=-1) if len(feats) > 1 else feats[0]
    flags = _lib.LP_FLAG_MASK_IMAGE | (_lib.LP_FLAG_REJECT_BEHIND if REJECT_BEHIND_CAMERA else 0) | \
        (_lib.LP_FLAG_BBOX_HALF_OPEN if BBOX_HALF_OPEN else 0) | (_lib.LP_FLAG_PLAIN_EPS if PLAIN_EPS else 0) | \
        (_lib.LP_FLAG_AFFINE_INTERP if AFFINE_INTERP else 0)
    cfg = functional.RenderConfig(
        verts=None, faces=None, cameras=None, proj=(1.0, 1.0, -1.0), H=int(height), W=int(width), flags=flags,
        multiplier=1000.0 if multiplier is None else float(multiplier), eps=1e-8 if eps is None else float(eps),
        fvi=functional._f32(face_vertices_image, device), fvz=functional._f32(face_vertices_z, device),
        valid=None if valid_faces is None else valid_faces.detach().to(device=device, dtype=torch.uint8).contiguous())
    image, _mask, face_idx, _bary, _depth = functional.render_face_features(ff.contiguous(), cfg)
    out = image.permute(0, 2, 3, 1)
    if is_list:
        out = tuple(torch.split(out, dims, dim=-1))
    return out, face_idx.long()


def dibr_rasterization(height, width, face_vertices_z, face_vertices_image, face_features, face_normals_z,
                       sigmainv=7000, boxlen=0.02, knum=30, multiplier=None, eps=None, rast_backend="cuda"):
    """Back-face rule ``face_normals_z > 0`` then :func:`rasterize`; the DIB-R soft mask (bound and never
    read by the reference, latent_paint_mesh/models/render.py:231) is replaced by the hard coverage."""
    feats, face_idx = rasterize(height, width, face_vertices_z, face_vertices_image, face_features,
                                valid_faces=face_normals_z > 0, multiplier=multiplier, eps=eps)
    return feats, (face_idx > -1).float(), face_idx


def texture_mapping(texture_coordinates, texture_maps, mode="nearest"):
    """(B,H,W,2), (B,C,T,T) → (B,H,W,C); gradients flow into ``texture_maps``."""
    B = texture_coordinates.shape[0]
    dims = texture_coordinates.shape[1:-1]
    C = texture_maps.shape[1]
    if texture_coordinates.dim() == 4:                                   # (B,H,W,2): the kernels' own layout
        out = functional.texture_map(texture_coordinates, texture_maps, mode)            # (B,C,H,W)
        return out.permute(0, 2, 3, 1)
    # flat coordinate lists: fold N into rows of 32 (padded with zeros, whose output is dropped and whose
    # upstream gradient is therefore zero), so the backward launches full warps on a legal grid
    uv = texture_coordinates.reshape(B, -1, 2)
    N = uv.shape[1]
    rows = (N + 31) // 32
    if rows * 32 != N:
        uv = torch.cat([uv, uv.new_zeros(B, rows * 32 - N, 2)], dim=1)
    out = functional.texture_map(uv.reshape(B, rows, 32, 2), texture_maps, mode)         # (B,C,rows,32)
    return out.reshape(B, C, rows * 32)[:, :, :N].permute(0, 2, 1).reshape(B, *dims, C)


SH_BAND1_AXES = (1, 2, 0)        # BASELINE.md decree 5


def spherical_harmonic_lighting(imnormal, lights):
    x, y, z = imnormal[..., 0], imnormal[..., 1], imnormal[..., 2]
    n = (x, y, z)
    bands = [0.28209479177 * torch.ones_like(x), 0.4886025119 * n[SH_BAND1_AXES[0]], 0.4886025119 * n[SH_BAND1_AXES[1]],
             0.4886025119 * n[SH_BAND1_AXES[2]], 1.09254843059 * (x * y), 1.09254843059 * (y * z),
             0.94617469575 * (z * z) - 0.31539156525, 0.77254840404 * (x * z), 0.38627420202 * (x * x - y * y)]
    L = lights.reshape(-1, 9)
    out = bands[0] * L[:, 0].reshape(-1, 1, 1)
    for i in range(1, 9):
        out = out + bands[i] * L[:, i].reshape(-1, 1, 1)
    return out


# ---------------------------------------------------------------- io.obj
def import_mesh(path, with_normals=False, with_materials=False):
    """→ object with ``vertices, faces, uvs, face_uvs_idx`` (reference mesh.py:11-24)."""
    return meshio.load_obj(path)


def import_off_mesh(path):
    """``kal.io.off.import_mesh`` (reference mesh.py:16-17)."""
    return meshio.load_off(path)


def make_module() -> types.ModuleType:
    kal = types.ModuleType("kaolin")
    render, cam, mesh = types.ModuleType("kaolin.render"), types.ModuleType("kaolin.render.camera"), \
        types.ModuleType("kaolin.render.mesh")
    ops, ops_mesh = types.ModuleType("kaolin.ops"), types.ModuleType("kaolin.ops.mesh")
    io, io_obj, io_off = types.ModuleType("kaolin.io"), types.ModuleType("kaolin.io.obj"), types.ModuleType("kaolin.io.off")
    cam.generate_perspective_projection = generate_perspective_projection
    cam.generate_transformation_matrix = generate_transformation_matrix
    mesh.prepare_vertices, mesh.rasterize, mesh.dibr_rasterization = prepare_vertices, rasterize, dibr_rasterization
    mesh.texture_mapping, mesh.spherical_harmonic_lighting = texture_mapping, spherical_harmonic_lighting
    ops_mesh.index_vertices_by_faces, ops_mesh.uniform_laplacian = index_vertices_by_faces, uniform_laplacian
    ops_mesh.uniform_laplacian_sparse = uniform_laplacian_sparse        # (extension: not a kaolin name)
    io_obj.import_mesh = import_mesh
    io_off.import_mesh = import_off_mesh
    render.camera, render.mesh, ops.mesh, io.obj, io.off = cam, mesh, ops_mesh, io_obj, io_off
    kal.render, kal.ops, kal.io = render, ops, io
    kal.__lp_b200__ = True
    return kal


def install() -> types.ModuleType:
    """Register the tree as ``kaolin`` in ``sys.modules`` (refuses to shadow a real kaolin)."""
    existing = sys.modules.get("kaolin")
    if existing is not None and not getattr(existing, "__lp_b200__", False) and not getattr(existing, "__oracle__", False):
        raise RuntimeError("a real kaolin is already imported; not shadowing it")
    kal = make_module()
    for name, mod in (("kaolin", kal), ("kaolin.render", kal.render), ("kaolin.render.camera", kal.render.camera),
                      ("kaolin.render.mesh", kal.render.mesh), ("kaolin.ops", kal.ops), ("kaolin.ops.mesh", kal.ops.mesh),
                      ("kaolin.io", kal.io), ("kaolin.io.obj", kal.io.obj), ("kaolin.io.off", kal.io.off)):
        sys.modules[name] = mod
    return kal
