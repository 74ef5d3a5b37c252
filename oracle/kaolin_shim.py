"""TEST INFRASTRUCTURE — CPU oracle: torch restatement of the kaolin functions on the
Latent-Paint render path.  Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` leg may import this; the product package never does.

The reference's per-pixel arithmetic lives in NVIDIA kaolin, installed un-pinned from git
master (reference ``setup.sh:3``, ``README.md:124-127``); it is not vendored in
``/root/reference``, not installed here and cannot be fetched.  This module therefore
restates the eight kaolin entry points the reference calls, following kaolin's documented
semantics, with the open points fixed by decree in ``BASELINE.md`` §4 / ``SURVEY.md``
Appendix A.  **Parity is unpinned at the kaolin boundary** (the reference has no tests,
golden vectors or fixtures); what *is* pinned: the reference's own Python glue
(``src/latent_paint/models/render.py``) is executed unmodified over this module by
``oracle/reference_glue.py`` to make ``tests/golden/*.npz``, and ``texture_mapping`` runs the
real ATen ``grid_sample`` kernel.

Call sites restated (reference file:line):
  camera.generate_perspective_projection   latent_paint/models/render.py:11; latent_paint_mesh/models/render.py:18-19
  camera.generate_transformation_matrix    latent_paint/models/render.py:30; latent_paint_mesh/models/render.py:54
  mesh.prepare_vertices                    latent_paint/models/render.py:39,56; latent_paint_mesh/models/render.py:194
  mesh.rasterize                           latent_paint/models/render.py:42,59
  mesh.dibr_rasterization                  latent_paint_mesh/models/render.py:231
  mesh.texture_mapping                     latent_paint/models/render.py:64; latent_paint_mesh/models/render.py:243
  mesh.spherical_harmonic_lighting         latent_paint_mesh/models/render.py:258
  ops.mesh.index_vertices_by_faces         latent_paint/models/textured_mesh.py:48; latent_paint_mesh/models/render.py:202

Arithmetic is fp32 with a fixed left-to-right expression order built from separate torch ops
(torch never fuses separate CPU ops, so there is no FMA contraction).
"""
from __future__ import annotations

import ctypes
import os
import sys
import types

import numpy as np
import torch

# ----------------------------------------------------------------------------- settings
#: rasterizer traversal: 'brute' (C, per pixel over all faces — normative), 'bbox' (C, per
#: face over its bounding box — fast), 'torch' (blocked dense torch — the "PyTorch CPU path"
#: timed as the reference arm).  All three give identical buffers (tests/test_oracle.py).
RASTER_IMPL = os.environ.get("LP_ORACLE_RASTER", "bbox")
#: decree 3 of BASELINE.md §4: faces whose interpolated depth is not < 0 never win a pixel
REJECT_BEHIND_CAMERA = True
#: the other open points of the decree as switches (defaults = the decree; a diff against real kaolin is a flag flip):
#: half-open bounding box (x0 < xmax, y0 < ymax), plain ``s + eps`` instead of ``s + copysign(eps, s)``, screen-space
#: (affine) instead of perspective-correct interpolation.  ``SH_BAND1_AXES`` below is the fourth.
BBOX_HALF_OPEN = False
PLAIN_EPS = False
AFFINE_INTERP = False


def decree_flags():
    return int(REJECT_BEHIND_CAMERA) | (2 if BBOX_HALF_OPEN else 0) | (4 if PLAIN_EPS else 0) | (8 if AFFINE_INTERP else 0)
DEFAULT_MULTIPLIER = 1000.0
DEFAULT_EPS = 1e-8

#: buffers of the most recent rasterize() call, for tests that run the reference's glue
#: (which returns neither face_idx nor the interpolated UVs)
LAST = {}

_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        here = os.path.dirname(os.path.abspath(__file__))
        sys.path.insert(0, here)
        try:
            import build as _oracle_build  # oracle/build.py
            path = _oracle_build.build()
        finally:
            sys.path.pop(0)
        lib = ctypes.CDLL(path)
        for name in ("lp_ref_rasterize_brute", "lp_ref_rasterize_bbox"):
            fn = getattr(lib, name)
            fn.restype = None
            fn.argtypes = [ctypes.c_int] * 5 + [ctypes.c_void_p] * 4 + [ctypes.c_float, ctypes.c_float, ctypes.c_int] \
                + [ctypes.c_void_p] * 4
        _LIB = lib
    return _LIB


# ----------------------------------------------------------------------------- camera
def generate_perspective_projection(fovyangle, ratio=1.0, dtype=torch.float):
    """(3,1) vector [1/(ratio·tan(fov/2)), 1/tan(fov/2), -1]."""
    tanfov = np.tan(fovyangle / 2.0)
    return torch.tensor([[1.0 / (ratio * tanfov)], [1.0 / tanfov], [-1]], dtype=dtype)


def _normalize3(v):
    n = torch.sqrt((v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]) + v[:, 2] * v[:, 2])
    return v / n[:, None]


def _cross3(a, b):
    return torch.stack([a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1],
                        a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2],
                        a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]], dim=1)


def generate_transformation_matrix(camera_position, look_at, camera_up_direction):
    """Look-at view matrix (B,4,3): rows 0-2 hold the camera axes x,y,z as columns, row 3 the
    translation -pos·R, so that ``[v,1] @ M`` is the camera-space point."""
    pos = camera_position.float()
    at = look_at.float().expand_as(pos)
    up = camera_up_direction.float().expand_as(pos)
    z = _normalize3(pos - at)
    x = _normalize3(_cross3(up, z))
    y = _cross3(z, x)
    rot = torch.stack([x, y, z], dim=2)                                     # (B,3,3)
    t = -((pos[:, 0:1] * rot[:, 0, :] + pos[:, 1:2] * rot[:, 1, :]) + pos[:, 2:3] * rot[:, 2, :])
    return torch.cat([rot, t[:, None, :]], dim=1)


# ----------------------------------------------------------------------------- ops.mesh
def index_vertices_by_faces(vertices_features, faces):
    """(B,V,K),(F,3) → (B,F,3,K) gather."""
    B, _, K = vertices_features.shape
    F = faces.shape[0]
    idx = faces.reshape(1, F * 3, 1).expand(B, F * 3, K)
    return torch.gather(vertices_features, 1, idx).reshape(B, F, 3, K)


def uniform_laplacian(num_vertices, faces):
    """Dense (V,V) uniform Laplacian (reference latent_paint_mesh/models/textured_mesh.py:60-71;
    off the hot path)."""
    adj = torch.zeros((num_vertices, num_vertices), dtype=torch.float32, device=faces.device)
    for a, b in ((0, 1), (1, 2), (2, 0)):
        adj[faces[:, a], faces[:, b]] = 1
        adj[faces[:, b], faces[:, a]] = 1
    deg = adj.sum(dim=1, keepdim=True).clamp(min=1)
    return adj / deg - torch.eye(num_vertices, device=faces.device)


# ----------------------------------------------------------------------------- render.mesh
def prepare_vertices(vertices, faces, camera_proj, camera_rot=None, camera_trans=None, camera_transform=None):
    """verts (V,3)|(B,V,3), faces (F,3), proj (3,1), transform (B,4,3) →
    face_vertices_camera (B,F,3,3), face_vertices_image (B,F,3,2), unit face normals (B,F,3)."""
    if camera_transform is None:
        raise NotImplementedError("the reference always passes camera_transform")
    v = vertices.float()
    if v.dim() == 2:
        v = v[None]
    M = camera_transform.float()
    vx, vy, vz = v[..., 0:1], v[..., 1:2], v[..., 2:3]
    # c_j = ((vx*M0j + vy*M1j) + vz*M2j) + M3j      — the reference's matmul with the order fixed
    cam = ((vx * M[:, None, 0, :] + vy * M[:, None, 1, :]) + vz * M[:, None, 2, :]) + M[:, None, 3, :]
    proj = camera_proj.float().reshape(1, 1, 3)
    pp = cam * proj
    img = pp[..., :2] / pp[..., 2:3]
    fvc = index_vertices_by_faces(cam, faces)
    fvi = index_vertices_by_faces(img, faces)
    e0 = fvc[:, :, 1] - fvc[:, :, 0]
    e1 = fvc[:, :, 2] - fvc[:, :, 0]
    n = torch.stack([e0[..., 1] * e1[..., 2] - e0[..., 2] * e1[..., 1],
                     e0[..., 2] * e1[..., 0] - e0[..., 0] * e1[..., 2],
                     e0[..., 0] * e1[..., 1] - e0[..., 1] * e1[..., 0]], dim=-1)
    ln = torch.sqrt((n[..., 0] * n[..., 0] + n[..., 1] * n[..., 1]) + n[..., 2] * n[..., 2])
    n = n / (ln[..., None] + 1e-10)
    return fvc, fvi, n


def _rasterize_c(H, W, fvz, fvi, valid, mult, eps, impl):
    B, F = fvz.shape[0], fvz.shape[1]
    fvz_c = fvz.detach().contiguous().float()
    fvi_c = fvi.detach().contiguous().float()
    face_idx = torch.empty((B, H, W), dtype=torch.int64)
    w = torch.empty((B, H, W, 3), dtype=torch.float32)
    depth = torch.empty((B, H, W), dtype=torch.float32)
    valid_c = valid.detach().contiguous().to(torch.uint8) if valid is not None else None
    fn = _lib().lp_ref_rasterize_brute if impl == "brute" else _lib().lp_ref_rasterize_bbox
    fn(B, F, H, W, 0, fvz_c.data_ptr(), fvi_c.data_ptr(), None,
       valid_c.data_ptr() if valid_c is not None else None,
       ctypes.c_float(mult), ctypes.c_float(eps), decree_flags(),
       None, face_idx.data_ptr(), w.data_ptr(), depth.data_ptr())
    return face_idx, w, depth


def _rasterize_torch(H, W, fvz, fvi, valid, mult, eps, block=32):
    """Blocked dense evaluation with torch ops only (all host threads) — the same expression
    tree as oracle/raster_ref.c:eval_face, so the buffers are identical."""
    B, F = fvz.shape[0], fvz.shape[1]
    f32 = torch.float32
    m = torch.tensor(mult, dtype=f32)
    X = (fvi.detach().float() * m)
    xa, ya, xb, yb, xc, yc = X[..., 0, 0], X[..., 0, 1], X[..., 1, 0], X[..., 1, 1], X[..., 2, 0], X[..., 2, 1]
    za, zb, zc = fvz[..., 0].detach().float(), fvz[..., 1].detach().float(), fvz[..., 2].detach().float()
    xmin, xmax = torch.minimum(torch.minimum(xa, xb), xc), torch.maximum(torch.maximum(xa, xb), xc)
    ymin, ymax = torch.minimum(torch.minimum(ya, yb), yc), torch.maximum(torch.maximum(ya, yb), yc)
    cols = (m / torch.tensor(float(W), dtype=f32)) * (2 * torch.arange(W) + 1 - W).to(f32)
    rows = (m / torch.tensor(float(H), dtype=f32)) * (H - 2 * torch.arange(H) - 1).to(f32)
    face_idx = torch.full((B, H, W), -1, dtype=torch.int64)
    wout = torch.zeros((B, H, W, 3), dtype=f32)
    depth = torch.zeros((B, H, W), dtype=f32)
    epsv = torch.tensor(eps, dtype=f32)
    for b in range(B):
        ok = torch.ones(F, dtype=torch.bool) if valid is None else valid[b].bool()
        for j0 in range(0, H, block):
            j1 = min(H, j0 + block)
            yhi, ylo = rows[j0], rows[j1 - 1]
            rowsel = ok & (ymax[b] >= ylo) & (ymin[b] <= yhi)
            for i0 in range(0, W, block):
                i1 = min(W, i0 + block)
                xlo, xhi = cols[i0], cols[i1 - 1]
                idx = torch.nonzero(rowsel & (xmax[b] >= xlo) & (xmin[b] <= xhi)).flatten()
                if idx.numel() == 0:
                    continue
                x0 = cols[i0:i1][None, :, None]
                y0 = rows[j0:j1][:, None, None]
                g = lambda t: t[b, idx][None, None, :]
                Xa, Ya, Xb, Yb, Xc, Yc = g(xa), g(ya), g(xb), g(yb), g(xc), g(yc)
                if BBOX_HALF_OPEN:
                    inbox = (g(xmin) <= x0) & (x0 < g(xmax)) & (g(ymin) <= y0) & (y0 < g(ymax))
                else:
                    inbox = (g(xmin) <= x0) & (x0 <= g(xmax)) & (g(ymin) <= y0) & (y0 <= g(ymax))
                w0 = (Xb - x0) * (Yc - y0) - (Yb - y0) * (Xc - x0)
                w1 = (Xc - x0) * (Ya - y0) - (Yc - y0) * (Xa - x0)
                w2 = (Xa - x0) * (Yb - y0) - (Ya - y0) * (Xb - x0)
                s = (w0 + w1) + w2
                s = s + (epsv if PLAIN_EPS else torch.copysign(epsv, s))
                w0, w1, w2 = w0 / s, w1 / s, w2 / s
                if AFFINE_INTERP:
                    z0 = (w0 * g(za) + w1 * g(zb)) + w2 * g(zc)
                else:
                    q = (w0 / g(za) + w1 / g(zb)) + w2 / g(zc)
                    z0 = 1.0 / q
                hit = inbox & (w0 >= 0) & (w1 >= 0) & (w2 >= 0)
                hit = hit & ((z0 < 0) if REJECT_BEHIND_CAMERA else (z0 == z0))
                key = torch.where(hit, z0, torch.full_like(z0, float("-inf")))
                best, arg = torch.max(key, dim=2)            # first maximal value = lowest face index
                anyhit = hit.any(dim=2)
                sel = arg[..., None]
                if AFFINE_INTERP:
                    pw = torch.stack([torch.gather(w0, 2, sel)[..., 0], torch.gather(w1, 2, sel)[..., 0],
                                      torch.gather(w2, 2, sel)[..., 0]], dim=-1)
                else:
                    pw = torch.stack([(torch.gather(w0, 2, sel)[..., 0] / za[b, idx][arg]) * best,
                                      (torch.gather(w1, 2, sel)[..., 0] / zb[b, idx][arg]) * best,
                                      (torch.gather(w2, 2, sel)[..., 0] / zc[b, idx][arg]) * best], dim=-1)
                face_idx[b, j0:j1, i0:i1] = torch.where(anyhit, idx[arg], torch.full_like(arg, -1))
                wout[b, j0:j1, i0:i1] = torch.where(anyhit[..., None], pw, torch.zeros_like(pw))
                depth[b, j0:j1, i0:i1] = torch.where(anyhit, best, torch.zeros_like(best))
    return face_idx, wout, depth


def rasterize_buffers(height, width, face_vertices_z, face_vertices_image, valid_faces=None,
                      multiplier=None, eps=None, impl=None):
    """Visibility buffers: face_idx (B,H,W) int64, perspective-correct weights w' (B,H,W,3),
    depth z0 (B,H,W).  Not differentiable (the reference never needs vertex gradients)."""
    mult = DEFAULT_MULTIPLIER if multiplier is None else float(multiplier)
    e = DEFAULT_EPS if eps is None else float(eps)
    impl = impl or RASTER_IMPL
    if impl == "torch":
        return _rasterize_torch(height, width, face_vertices_z, face_vertices_image, valid_faces, mult, e)
    return _rasterize_c(height, width, face_vertices_z, face_vertices_image, valid_faces, mult, e, impl)


def _interpolate(face_idx, w, face_features):
    """feat = (w'0·fa + w'1·fb) + w'2·fc on covered pixels, 0 elsewhere; differentiable in
    ``face_features`` (that is kaolin's rasterize backward into the face features)."""
    B, H, W = face_idx.shape
    F, D = face_features.shape[1], face_features.shape[3]
    covered = face_idx >= 0
    flat = (face_idx.clamp(min=0) + (torch.arange(B)[:, None, None] * F)).reshape(-1)
    ff = face_features.reshape(B * F, 3, D)[flat].reshape(B, H, W, 3, D)
    out = (w[..., 0:1] * ff[..., 0, :] + w[..., 1:2] * ff[..., 1, :]) + w[..., 2:3] * ff[..., 2, :]
    return out * covered[..., None].to(out.dtype)


def rasterize(height, width, face_vertices_z, face_vertices_image, face_features, valid_faces=None,
              multiplier=None, eps=None, backend="cuda"):
    """→ (interpolated_features (B,H,W,D) or a tuple of them, face_idx (B,H,W) int64)."""
    face_idx, w, depth = rasterize_buffers(height, width, face_vertices_z, face_vertices_image, valid_faces,
                                           multiplier, eps)
    if isinstance(face_features, (list, tuple)):
        feats = tuple(_interpolate(face_idx, w, f) for f in face_features)
    else:
        feats = _interpolate(face_idx, w, face_features)
    LAST.clear()
    LAST.update(face_idx=face_idx, bary=w, depth=depth,
                features=feats[0].detach() if isinstance(feats, tuple) else feats.detach())
    return feats, face_idx


def dibr_rasterization(height, width, face_vertices_z, face_vertices_image, face_features, face_normals_z,
                       sigmainv=7000, boxlen=0.02, knum=30, multiplier=None, eps=None, rast_backend="cuda"):
    """Back-face rule ``face_normals_z > 0`` then :func:`rasterize`.  The DIB-R soft mask is
    not computed (the reference binds it and never reads it, latent_paint_mesh/models/render.py:231);
    the hard coverage is returned in its place."""
    valid = face_normals_z > 0
    feats, face_idx = rasterize(height, width, face_vertices_z, face_vertices_image, face_features,
                                valid_faces=valid, multiplier=multiplier, eps=eps)
    soft_mask = (face_idx > -1).float()
    return feats, soft_mask, face_idx


def texture_mapping(texture_coordinates, texture_maps, mode="nearest"):
    """clamp → [-1,1] → flip v → the real ATen ``grid_sample(align_corners=False, border)``."""
    B = texture_coordinates.shape[0]
    C = texture_maps.shape[1]
    dims = texture_coordinates.shape[1:-1]
    g = texture_coordinates.reshape(B, -1, 1, 2)
    g = torch.clamp(g, 0.0, 1.0)
    g = g * 2 - 1
    g = torch.stack([g[..., 0], -g[..., 1]], dim=-1)
    out = torch.nn.functional.grid_sample(texture_maps, g, mode=mode, align_corners=False, padding_mode="border")
    return out.permute(0, 2, 3, 1).reshape(B, *dims, C)


#: real SH basis used for lighting (BASELINE.md decree 5).  Band-1 axis
#: order (y, z, x) is the decree; kaolin's own order could not be checked here.
SH_BAND1_AXES = (1, 2, 0)


def spherical_harmonic_lighting(imnormal, lights):
    x, y, z = imnormal[..., 0], imnormal[..., 1], imnormal[..., 2]
    n = (x, y, z)
    bands = [0.28209479177 * torch.ones_like(x),
             0.4886025119 * n[SH_BAND1_AXES[0]],
             0.4886025119 * n[SH_BAND1_AXES[1]],
             0.4886025119 * n[SH_BAND1_AXES[2]],
             1.09254843059 * (x * y),
             1.09254843059 * (y * z),
             0.94617469575 * (z * z) - 0.31539156525,
             0.77254840404 * (x * z),
             0.38627420202 * (x * x - y * y)]
    L = lights.reshape(-1, 9)
    out = bands[0] * L[:, 0].reshape(-1, 1, 1)
    for i in range(1, 9):
        out = out + bands[i] * L[:, i].reshape(-1, 1, 1)
    return out


# ----------------------------------------------------------------------------- module tree
def make_module() -> types.ModuleType:
    """Build a module tree named ``kaolin`` exposing exactly what the reference imports."""
    kal = types.ModuleType("kaolin")
    render, camera, mesh = types.ModuleType("kaolin.render"), types.ModuleType("kaolin.render.camera"), \
        types.ModuleType("kaolin.render.mesh")
    ops, ops_mesh = types.ModuleType("kaolin.ops"), types.ModuleType("kaolin.ops.mesh")
    camera.generate_perspective_projection = generate_perspective_projection
    camera.generate_transformation_matrix = generate_transformation_matrix
    mesh.prepare_vertices = prepare_vertices
    mesh.rasterize = rasterize
    mesh.dibr_rasterization = dibr_rasterization
    mesh.texture_mapping = texture_mapping
    mesh.spherical_harmonic_lighting = spherical_harmonic_lighting
    ops_mesh.index_vertices_by_faces = index_vertices_by_faces
    ops_mesh.uniform_laplacian = uniform_laplacian
    render.camera, render.mesh, ops.mesh = camera, mesh, ops_mesh
    kal.render, kal.ops = render, ops
    kal.__oracle__ = True
    return kal


def install() -> types.ModuleType:
    """Inject the tree into ``sys.modules`` so ``import kaolin as kal`` in the reference's own
    files resolves to this oracle (tests / golden generation only)."""
    kal = make_module()
    sys.modules["kaolin"] = kal
    sys.modules["kaolin.render"] = kal.render
    sys.modules["kaolin.render.camera"] = kal.render.camera
    sys.modules["kaolin.render.mesh"] = kal.render.mesh
    sys.modules["kaolin.ops"] = kal.ops
    sys.modules["kaolin.ops.mesh"] = kal.ops.mesh
    return kal
