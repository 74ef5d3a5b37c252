"""Mesh input for the renderer: OBJ reader, normalisation, UV provisioning.

Replaces what the reference gets from ``kal.io.obj.import_mesh`` and
``Mesh.normalize_mesh`` (reference ``src/latent_paint/models/mesh.py:6-48``) and the UV
source selection of ``TexturedMeshModel.init_texture_map``
(reference ``src/latent_paint/models/textured_mesh.py:81-109``).  kaolin and xatlas are
not available, so meshes without a complete UV set get the deterministic per-face grid
atlas declared in SURVEY.md §8(d) / BASELINE.md instead of an xatlas parametrisation.

Everything here is host-side numpy/torch plumbing that runs once per mesh.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np
import torch


@dataclass
class MeshData:
    """What ``kal.io.obj.import_mesh`` returns, reduced to the fields the reference reads
    (``mesh.py:21-24``): ``vertices (V,3) f32``, ``faces (F,3) i64``, ``uvs (Nt,2) f32``,
    ``face_uvs_idx (F,3) i64`` with -1 where a face has no ``vt`` (the reference tests
    ``ft.min() > -1``, ``textured_mesh.py:84-85``)."""

    vertices: torch.Tensor
    faces: torch.Tensor
    uvs: torch.Tensor
    face_uvs_idx: torch.Tensor


def load_obj(path: str) -> MeshData:
    """Plain OBJ reader: ``v``, ``vt`` and triangular ``f`` records with 1-based
    ``v``, ``v/vt``, ``v//vn`` or ``v/vt/vn`` corners.  ``vt`` is always parsed (the
    reference only gets UVs through kaolin's ``with_materials=True`` route, mesh.py:11-14)."""
    verts, uvs, faces, fuv = [], [], [], []
    with open(path, "r") as fh:
        for line in fh:
            if line.startswith("v "):
                p = line.split()
                verts.append((float(p[1]), float(p[2]), float(p[3])))
            elif line.startswith("vt "):
                p = line.split()
                uvs.append((float(p[1]), float(p[2])))
            elif line.startswith("f "):
                p = line.split()[1:]
                if len(p) != 3:
                    raise ValueError(f"{path}: only triangular faces are supported, got {len(p)} corners")
                vi, ti = [], []
                for c in p:
                    s = c.split("/")
                    vi.append(int(s[0]) - 1)
                    ti.append(int(s[1]) - 1 if len(s) > 1 and s[1] != "" else -1)
                faces.append(vi)
                fuv.append(ti)
    v = torch.tensor(np.asarray(verts, dtype=np.float32).reshape(-1, 3))
    f = torch.tensor(np.asarray(faces, dtype=np.int64).reshape(-1, 3))
    vt = torch.tensor(np.asarray(uvs, dtype=np.float32).reshape(-1, 2))
    ft = torch.tensor(np.asarray(fuv, dtype=np.int64).reshape(-1, 3))
    return MeshData(v, f, vt, ft)


def load_off(path: str) -> MeshData:
    """Object File Format reader (``kal.io.off.import_mesh``, reference mesh.py:16-17): ``OFF`` header,
    ``nv nf ne`` counts, vertex rows, then ``3 i j k`` face rows (0-based).  OFF carries no UVs."""
    with open(path, "r") as fh:
        tok = [t for line in fh for t in line.split("#", 1)[0].split()]
    if not tok or not tok[0].upper().startswith("OFF"):
        raise ValueError(f"{path}: not an OFF file")
    head = tok[0][3:]                      # "OFF3 4 0" style headers glue the first count to the tag
    rest = ([head] if head else []) + tok[1:]
    nv, nf = int(rest[0]), int(rest[1])
    pos = 3
    verts = np.asarray(rest[pos:pos + 3 * nv], dtype=np.float32).reshape(nv, 3)
    pos += 3 * nv
    faces = []
    for _ in range(nf):
        n = int(rest[pos])
        if n != 3:
            raise ValueError(f"{path}: only triangular faces are supported, got {n} corners")
        faces.append([int(rest[pos + 1]), int(rest[pos + 2]), int(rest[pos + 3])])
        pos += 1 + n
    f = torch.tensor(np.asarray(faces, dtype=np.int64).reshape(-1, 3))
    return MeshData(torch.tensor(verts), f, torch.zeros((0, 2)), torch.full((f.shape[0], 3), -1, dtype=torch.int64))


def save_npz(mesh: MeshData, path: str) -> None:
    np.savez_compressed(path, vertices=mesh.vertices.numpy(), faces=mesh.faces.numpy().astype(np.int32),
                        uvs=mesh.uvs.numpy(), face_uvs_idx=mesh.face_uvs_idx.numpy().astype(np.int32))


def load_npz(path: str) -> MeshData:
    z = np.load(path)
    return MeshData(torch.tensor(z["vertices"]), torch.tensor(z["faces"].astype(np.int64)),
                    torch.tensor(z["uvs"].reshape(-1, 2)), torch.tensor(z["face_uvs_idx"].astype(np.int64)))


def find_shape(name: str, shapes_dir: str | None = None) -> MeshData:
    """Locate ``shapes/<name>.obj``.  Search order: explicit dir, ``$LP_SHAPES_DIR``, the
    reference checkout (this container only), then the packed arrays committed under
    ``tests/golden/meshes`` (what the GPU box sees; written by ``tests/golden/make_golden.py``)."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cands = [shapes_dir, os.environ.get("LP_SHAPES_DIR"), "/root/reference/shapes"]
    for d in cands:
        if d and os.path.isfile(os.path.join(d, name + ".obj")):
            return load_obj(os.path.join(d, name + ".obj"))
    packed = os.path.join(here, "tests", "golden", "meshes", name + ".npz")
    if os.path.isfile(packed):
        return load_npz(packed)
    raise FileNotFoundError(f"mesh '{name}' not found in {cands} nor {packed}")


def normalize_vertices(vertices: torch.Tensor, target_scale: float = 1.0, dy: float = 0.0) -> torch.Tensor:
    """``Mesh.normalize_mesh`` (reference mesh.py:37-48): centre by the vertex mean, divide
    by the largest vertex norm, scale, lift by ``dy``."""
    verts = vertices - vertices.mean(dim=0)
    verts = verts / torch.max(torch.norm(verts, p=2, dim=1))
    verts = verts * target_scale
    verts[:, 1] += dy
    return verts


def grid_atlas_uvs(num_faces: int) -> tuple[torch.Tensor, torch.Tensor]:
    """Deterministic per-face atlas standing in for xatlas: n =
    ceil(sqrt(F/2)) cells per side; faces 2k / 2k+1 are the lower-left / upper-right
    triangles of cell k, inset by 5 % of a cell.  Returns ``(vt (3F,2), ft (F,3))``."""
    n = int(np.ceil(np.sqrt(num_faces / 2.0)))
    k = np.arange(num_faces) // 2
    cx = (k % n).astype(np.float64)
    cy = (k // n).astype(np.float64)
    inset = 0.05
    lo, hi = inset, 1.0 - inset
    # lower-left triangle: (lo,lo) (hi,lo) (lo,hi); upper-right: (hi,hi) (lo,hi) (hi,lo)
    odd = (np.arange(num_faces) % 2).astype(bool)
    tri = np.empty((num_faces, 3, 2), dtype=np.float64)
    tri[~odd] = np.array([[lo, lo], [hi, lo], [lo, hi]])
    tri[odd] = np.array([[hi, hi], [lo, hi], [hi, lo]])
    tri[:, :, 0] = (tri[:, :, 0] + cx[:, None]) / n
    tri[:, :, 1] = (tri[:, :, 1] + cy[:, None]) / n
    vt = torch.tensor(tri.reshape(-1, 2).astype(np.float32))
    ft = torch.arange(num_faces * 3, dtype=torch.int64).reshape(num_faces, 3)
    return vt, ft


def face_uv_attributes(mesh: MeshData) -> torch.Tensor:
    """``kal.ops.mesh.index_vertices_by_faces(vt[None], ft)`` → ``(1,F,3,2)`` exactly as
    the reference builds ``face_attributes`` (textured_mesh.py:48-50); falls back to the
    grid atlas when any face lacks UVs (the reference would call xatlas there,
    textured_mesh.py:91-108)."""
    vt, ft = mesh.uvs, mesh.face_uvs_idx
    if vt is None or ft is None or vt.shape[0] == 0 or ft.numel() == 0 or int(ft.min()) < 0:
        vt, ft = grid_atlas_uvs(mesh.faces.shape[0])
    return vt[ft.reshape(-1)].reshape(1, -1, 3, 2).contiguous()


def provision_uvs(mesh, cache_dir=None, save_cache: bool = True):
    """UV source selection of ``TexturedMeshModel.init_texture_map`` (reference
    ``src/latent_paint/models/textured_mesh.py:81-109``, same in ``latent_paint_mesh``), returning ``(vt (Nt,2) f32,
    ft (F,3))``:

    1. the mesh's own UVs when every face has them (``vt.shape[0] > 0 and ft.min() > -1``, :84-87);
    2. else the cache ``<cache_dir>/vt.pth`` + ``ft.pth`` when both exist (:88-90) — the reference's on-disk format:
       ``torch.save`` of a CPU float32 ``(Nt,2)`` tensor and a CPU int32 ``(F,3)`` tensor (:103-107);
    3. else a fresh parametrisation, written to the cache in that format.  The reference runs xatlas here
       (:91-102); xatlas is not available, so this is the deterministic per-face grid atlas.

    ``mesh`` is a ``MeshData`` or anything with ``vt``/``ft`` (the reference ``Mesh``) or ``uvs``/``face_uvs_idx``."""
    vt = getattr(mesh, "vt", None) if hasattr(mesh, "vt") else getattr(mesh, "uvs", None)
    ft = getattr(mesh, "ft", None) if hasattr(mesh, "ft") else getattr(mesh, "face_uvs_idx", None)
    if vt is not None and ft is not None and vt.shape[0] > 0 and ft.numel() > 0 and int(ft.min()) > -1:
        return vt, ft
    vt_cache = os.path.join(str(cache_dir), "vt.pth") if cache_dir is not None else None
    ft_cache = os.path.join(str(cache_dir), "ft.pth") if cache_dir is not None else None
    if vt_cache and os.path.isfile(vt_cache) and os.path.isfile(ft_cache):
        return torch.load(vt_cache), torch.load(ft_cache)
    vt, ft = grid_atlas_uvs(int(mesh.faces.shape[0]))
    ft = ft.int()
    if vt_cache and save_cache:
        os.makedirs(str(cache_dir), exist_ok=True)
        torch.save(vt.cpu(), vt_cache)
        torch.save(ft.cpu(), ft_cache)
    return vt, ft


class Mesh:
    """Mirror of the reference ``Mesh`` (``src/latent_paint/models/mesh.py:6-48``): ``.vertices``, ``.faces`` on
    ``device``, ``.vt`` / ``.ft`` as loaded (``ft`` is -1 where a face has no UVs), ``standardize_mesh`` and
    ``normalize_mesh`` with the reference's arithmetic.  Reads ``.obj`` and ``.off`` without kaolin."""

    def __init__(self, obj_path, device):
        path = str(obj_path)
        if ".obj" in path:
            mesh = load_obj(path)
        elif ".off" in path:
            mesh = load_off(path)
        else:
            raise ValueError(f"{obj_path} extension not implemented in mesh reader.")
        self.vertices = mesh.vertices.to(device)
        self.faces = mesh.faces.to(device)
        self.ft = mesh.face_uvs_idx
        self.vt = mesh.uvs

    def standardize_mesh(self, inplace=False):
        import copy
        mesh = self if inplace else copy.deepcopy(self)
        verts = mesh.vertices
        verts = verts - verts.mean(dim=0)
        mesh.vertices = verts / torch.std(torch.norm(verts, p=2, dim=1))
        return mesh

    def normalize_mesh(self, inplace=False, target_scale=1, dy=0):
        import copy
        mesh = self if inplace else copy.deepcopy(self)
        mesh.vertices = normalize_vertices(mesh.vertices, target_scale, dy)
        return mesh


def subdivide(mesh: MeshData, levels: int, project_to_sphere: bool = True) -> MeshData:
    """Midpoint subdivision (each triangle → 4) used to build the config-4 stress mesh
    from ``sphere.obj`` (1280·4^5 = 1 310 720 faces).  UVs are subdivided per face corner."""
    v = mesh.vertices.numpy().astype(np.float64)
    f = mesh.faces.numpy()
    has_uv = mesh.uvs.shape[0] > 0 and int(mesh.face_uvs_idx.min()) >= 0
    fuv = mesh.uvs.numpy().astype(np.float64)[mesh.face_uvs_idx.numpy()] if has_uv else None  # (F,3,2)
    for _ in range(levels):
        nv = v.shape[0]
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], axis=0)
        es = np.sort(e, axis=1)
        key = es[:, 0].astype(np.int64) * nv + es[:, 1]
        uniq, inv = np.unique(key, return_inverse=True)
        mid = 0.5 * (v[uniq // nv] + v[uniq % nv])
        if project_to_sphere:
            mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        v = np.concatenate([v, mid], axis=0)
        nf = f.shape[0]
        m01, m12, m20 = nv + inv[:nf], nv + inv[nf:2 * nf], nv + inv[2 * nf:]
        a, b, c = f[:, 0], f[:, 1], f[:, 2]
        f = np.concatenate([np.stack([a, m01, m20], 1), np.stack([m01, b, m12], 1),
                            np.stack([m20, m12, c], 1), np.stack([m01, m12, m20], 1)], axis=0)
        if fuv is not None:
            ua, ub, uc = fuv[:, 0], fuv[:, 1], fuv[:, 2]
            u01, u12, u20 = 0.5 * (ua + ub), 0.5 * (ub + uc), 0.5 * (uc + ua)
            fuv = np.concatenate([np.stack([ua, u01, u20], 1), np.stack([u01, ub, u12], 1),
                                  np.stack([u20, u12, uc], 1), np.stack([u01, u12, u20], 1)], axis=0)
    F = f.shape[0]
    if fuv is not None:
        vt = torch.tensor(fuv.reshape(-1, 2).astype(np.float32))
        ft = torch.arange(F * 3, dtype=torch.int64).reshape(F, 3)
    else:
        vt, ft = torch.zeros((0, 2)), torch.full((F, 3), -1, dtype=torch.int64)
    return MeshData(torch.tensor(v.astype(np.float32)), torch.tensor(f.astype(np.int64)), vt, ft)
