"""Shared helpers for the parity tests: seeded scenes in the reference's own distributions
 and the tolerance BASELINE.json states for floating point."""
import os

import numpy as np
import torch

import latent_nerf_test_b200 as lp

RTOL, ATOL = 1e-4, 1e-5          # north_star: rendered pixels and texture gradients
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def scene(shape, scale, dy, subdivide=0):
    m = lp.meshio.find_shape(shape)
    if subdivide:
        m = lp.meshio.subdivide(m, subdivide)
    verts = lp.meshio.normalize_vertices(m.vertices, scale, dy)
    return verts, m.faces, lp.meshio.face_uv_attributes(m)


def scene_raw(shape):
    """Un-normalised mesh (the environment sphere is used at its own radius-20 scale, reference textured_mesh.py:53-54)."""
    m = lp.meshio.find_shape(shape)
    return m.vertices, m.faces, m


def rnd(shape, seed, scale=1.0):
    return scale * torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def latent_paint_views(B, seed=0):
    """radius, theta, phi drawn in the reference's order (latent_paint views_dataset.py:16-18);
    theta from 15° to avoid the up‖view singularity."""
    g = torch.Generator().manual_seed(seed)
    radius = torch.rand(B, generator=g) * 0.5 + 1.0
    theta = torch.deg2rad(torch.rand(B, generator=g) * 135.0 + 15.0)
    phi = torch.deg2rad(torch.rand(B, generator=g) * 360.0)
    return radius, theta, phi


def mesh_views(B, seed=0):
    """latent_paint_mesh train_config.py:18-22 ranges."""
    g = torch.Generator().manual_seed(seed)
    radius = torch.rand(B, generator=g) * 1.0 + 1.4
    theta = torch.deg2rad(torch.rand(B, generator=g) * 50.0 + 60.0)
    phi = torch.deg2rad(torch.rand(B, generator=g) * 360.0)
    return radius, theta, phi


def assert_close(a, b, what, rtol=RTOL, atol=ATOL):
    a = torch.as_tensor(a).detach().cpu().float()
    b = torch.as_tensor(b).detach().cpu().float()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    bad = err > tol
    assert not bad.any(), f"{what}: {int(bad.sum())} of {bad.numel()} elements off, max abs err {float(err.max()):.3e}"
