"""latent-nerf-test_b200 — B200-native Latent-Paint differentiable mesh renderer.

One hot path of chacorp/latent-nerf-test, rebuilt for sm_100a behind the reference's own
renderer interface:

    from latent_nerf_test_b200 import LatentPaintRenderer, LatentPaintMeshRenderer

``LatentPaintRenderer`` replaces ``src.latent_paint.models.render.Renderer``,
``LatentPaintMeshRenderer`` replaces ``src.latent_paint_mesh.models.render.Renderer``
(see INTEGRATION.md).  The directory name carries a hyphen, so the importable alias
``latent_nerf_test_b200`` is provided by ``latent_nerf_test_b200.py`` at the repo root.
"""
from . import _lib, camera, functional, kaolin_compat, meshio, optim, textured_mesh  # noqa: F401
from .render import Renderer as LatentPaintRenderer  # noqa: F401
from .render_mesh import Renderer as LatentPaintMeshRenderer  # noqa: F401

__all__ = ["LatentPaintRenderer", "LatentPaintMeshRenderer", "camera", "functional", "kaolin_compat", "meshio", "optim", "textured_mesh", "_lib"]
