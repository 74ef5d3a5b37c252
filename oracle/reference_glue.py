"""TEST INFRASTRUCTURE — runs the reference's OWN renderer glue on CPU over the oracle's kaolin
restatement.  Works only where /root/reference exists (this container); used by
``tests/golden/make_golden.py`` to freeze fixtures and by the CPU tests that pin
``oracle/renderer_ref.py`` (the travelling mirror) against the real reference files.

  * ``src/latent_paint/models/render.py`` is imported **unmodified**.
  * ``src/latent_paint_mesh/models/render.py`` hard-codes ``device='cuda'`` (lines 141, 226,
    325); its source text is patched in memory ('cuda' → 'cpu' in those literals) and exec'd.
    Nothing is written anywhere.
"""
from __future__ import annotations

import importlib.util
import os
import types

from . import kaolin_shim

REFERENCE_ROOT = os.environ.get("LP_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src/latent_paint/models/render.py"))


def load_latent_paint_renderer():
    """→ the reference class ``src.latent_paint.models.render.Renderer`` (unmodified file)."""
    kaolin_shim.install()
    path = os.path.join(REFERENCE_ROOT, "src/latent_paint/models/render.py")
    spec = importlib.util.spec_from_file_location("_ref_latent_paint_render", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.Renderer


def load_latent_paint_mesh_renderer():
    """→ the reference class ``src.latent_paint_mesh.models.render.Renderer`` with its three
    ``device='cuda'`` literals redirected to CPU (in memory only)."""
    kaolin_shim.install()
    path = os.path.join(REFERENCE_ROOT, "src/latent_paint_mesh/models/render.py")
    with open(path, "r") as fh:
        src = fh.read()
    src = src.replace("device='cuda'", "device='cpu'")
    mod = types.ModuleType("_ref_latent_paint_mesh_render")
    mod.__file__ = path
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod.Renderer
