"""Freeze golden vectors for the render path.  RUN IN THE BUILD CONTAINER ONLY (needs
/root/reference):   python tests/golden/make_golden.py

What is frozen
  tests/golden/meshes/<name>.npz   packed arrays of reference shapes/<name>.obj (inputs only; the GPU
                                   box has no /root/reference)
  tests/golden/<case>.npz          inputs + outputs of the reference's OWN renderer glue
                                   (src/latent_paint/models/render.py unmodified;
                                   src/latent_paint_mesh/models/render.py with its 'cuda' literals
                                   redirected in memory) executed on CPU over oracle/kaolin_shim.py with
                                   the normative brute-force rasterizer.
kaolin itself is not available, so these vectors pin the reference glue + the decreed kaolin
restatement + the real ATen grid_sample — not kaolin's own kernels.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import latent_nerf_test_b200 as lp  # noqa: E402  (meshio only; no kernels are run here)
from oracle import kaolin_shim, reference_glue  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SHAPES = ["blub", "nascar", "teddy", "sphere", "env_sphere"]


def rnd(shape, seed, scale=1.0):
    return scale * torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def pack_meshes():
    os.makedirs(os.path.join(OUT, "meshes"), exist_ok=True)
    for name in SHAPES:
        m = lp.meshio.load_obj(os.path.join(reference_glue.REFERENCE_ROOT, "shapes", name + ".obj"))
        lp.meshio.save_npz(m, os.path.join(OUT, "meshes", name + ".npz"))
        print("packed", name, tuple(m.vertices.shape), tuple(m.faces.shape))


def save(case, **arrays):
    arrays = {k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrays.items()}
    np.savez_compressed(os.path.join(OUT, case + ".npz"), **arrays)
    print("golden", case, {k: v.shape for k, v in arrays.items() if hasattr(v, "shape") and v.ndim > 0})


def latent_paint_texture(case, shape, scale, dy, dims, T, C, mode, elev, azim, radius, white):
    kaolin_shim.RASTER_IMPL = "brute"
    R = reference_glue.load_latent_paint_renderer()
    m = lp.meshio.find_shape(shape)
    verts = lp.meshio.normalize_vertices(m.vertices, scale, dy)
    uv = lp.meshio.face_uv_attributes(m)
    tex = rnd((1, C, T, T), 1, 0.4).requires_grad_(True)
    r = R("cpu", dim=dims, interpolation_mode=mode)
    image, mask = r.render_single_view_texture(verts, m.faces, uv, tex, elev=elev, azim=azim, radius=radius,
                                               look_at_height=dy, white_background=white)
    g = rnd(tuple(image.shape), 2)
    image.backward(g)
    save(case, shape=shape, scale=scale, dy=dy, dims=np.array(dims), mode=mode, elev=elev, azim=azim, radius=radius,
         white=white, texture=tex, grad_image=g, image=image.contiguous(), mask=mask.contiguous(),
         face_idx=kaolin_shim.LAST["face_idx"].to(torch.int32), uv=kaolin_shim.LAST["features"],
         grad_texture=tex.grad)


def latent_paint_colors(case, dims, elev, azim, radius):
    kaolin_shim.RASTER_IMPL = "brute"
    R = reference_glue.load_latent_paint_renderer()
    m = lp.meshio.find_shape("env_sphere")

    class _Mesh:
        vertices, faces = m.vertices, m.faces
    colors = rnd((1, m.faces.shape[0], 3, 4), 3).requires_grad_(True)
    r = R("cpu", dim=dims)
    image, mask = r.render_single_view(_Mesh, colors, elev=elev, azim=azim, radius=radius, look_at_height=0.25)
    g = rnd(tuple(image.shape), 4)
    image.backward(g)
    save(case, dims=np.array(dims), elev=elev, azim=azim, radius=radius, colors=colors, grad_image=g,
         image=image.contiguous(), mask=mask.contiguous(), face_idx=kaolin_shim.LAST["face_idx"].to(torch.int32),
         grad_colors=colors.grad)


def mesh_texture(case, shape, dims, T, C, B, is_body, white, radius_float=None):
    kaolin_shim.RASTER_IMPL = "brute"
    R = reference_glue.load_latent_paint_mesh_renderer()
    m = lp.meshio.find_shape(shape)
    verts = lp.meshio.normalize_vertices(m.vertices, 1.0, 0.0)
    uv = lp.meshio.face_uv_attributes(m)
    tex = rnd((1, C, T, T), 1, 0.4).requires_grad_(True)
    g0 = torch.Generator().manual_seed(0)
    radius = torch.rand(B, generator=g0) * 1.0 + 1.4
    elev = torch.deg2rad(torch.rand(B, generator=g0) * 50.0 + 60.0)
    azim = torch.deg2rad(torch.rand(B, generator=g0) * 360.0)
    r = R("cpu", dim=dims, interpolation_mode="bilinear")
    rad = radius if radius_float is None else radius_float
    image, mask, normals, lighting = r.render_single_view_texture(verts, m.faces, uv, tex, elev, azim, rad, dims=dims,
                                                                  white_background=white, is_body=is_body)
    g = rnd(tuple(image.shape), 2)
    image.backward(g)
    save(case, shape=shape, dims=np.array(dims), is_body=is_body, white=white, elev=elev, azim=azim,
         radius=radius if radius_float is None else np.float32(radius_float), texture=tex, grad_image=g,
         image=image.contiguous(), mask=mask.contiguous(), normals=normals.contiguous(), lighting=lighting.contiguous(),
         face_idx=kaolin_shim.LAST["face_idx"].to(torch.int32), grad_texture=tex.grad)


def latent_paint_render_train(case, dims, elev, azim, radius):
    """TexturedMeshModel.render_train (reference src/latent_paint/models/textured_mesh.py:187-220): the model class
    cannot be imported (hard-coded .cuda(), xatlas), so its render lines are restated in
    oracle/renderer_ref.render_train_ref and run here over the reference's REAL Renderer class."""
    from oracle import renderer_ref
    kaolin_shim.RASTER_IMPL = "brute"
    R = reference_glue.load_latent_paint_renderer()
    m, env = lp.meshio.find_shape("blub"), lp.meshio.find_shape("env_sphere")
    verts = lp.meshio.normalize_vertices(m.vertices, 0.6, 0.25)
    uv = lp.meshio.face_uv_attributes(m)
    tex = rnd((1, 4, 128, 128), 1, 0.4).requires_grad_(True)
    colors = torch.rand(1, env.faces.shape[0], 3, 4, generator=torch.Generator().manual_seed(3)).requires_grad_(True)
    r = R("cpu", dim=dims, interpolation_mode="bilinear")
    out = renderer_ref.render_train_ref(r, verts, m.faces, uv, tex, env.vertices, env.faces, colors, elev, azim, radius, dy=0.25)
    g = rnd(tuple(out["image"].shape), 4)
    out["image"].backward(g)
    save(case, dims=np.array(dims), elev=elev, azim=azim, radius=radius, texture=tex, colors=colors, grad_image=g,
         image=out["image"].contiguous(), mask=out["mask"].contiguous(), background=out["background"].contiguous(),
         foreground=out["foreground"].contiguous(), grad_texture=tex.grad, grad_colors=colors.grad)


if __name__ == "__main__":
    assert reference_glue.available(), "needs /root/reference"
    if len(sys.argv) > 1 and sys.argv[1] == "render_train":      # only the fixture added after the first freeze
        latent_paint_render_train("lp_render_train_blub", (64, 64), 1.0, 0.7, 1.25)
        sys.exit(0)
    pack_meshes()
    latent_paint_texture("lp_blub_nearest", "blub", 0.6, 0.25, (64, 64), 128, 4, "nearest", 1.0, 0.7, 1.25, False)
    latent_paint_texture("lp_blub_bilinear_white", "blub", 0.6, 0.25, (96, 96), 128, 3, "bilinear", 0.6, 2.1, 1.1, True)
    latent_paint_colors("lp_env_sphere_colors", (64, 64), 1.0, 0.7, 1.25)
    mesh_texture("mesh_sphere_body_b3", "sphere", (64, 64), 32, 4, 3, True, False)
    mesh_texture("mesh_teddy_head_white", "teddy", (48, 48), 64, 3, 2, False, True, radius_float=2.0)
    latent_paint_render_train("lp_render_train_blub", (64, 64), 1.0, 0.7, 1.25)
