"""CPU tests of the oracle itself: the three rasterizer traversals agree, the travelling mirror
reproduces the frozen reference-glue vectors, analytic known answers hold."""
import numpy as np
import pytest
import torch

from oracle import kaolin_shim as kal
from oracle import reference_glue, renderer_ref
from tests.common import assert_close, load_golden, rnd, scene, scene_raw


def _random_soup(B, F, seed, behind=False):
    g = torch.Generator().manual_seed(seed)
    centre = torch.rand(B, F, 1, 2, generator=g) * 2.4 - 1.2
    fvi = centre + (torch.rand(B, F, 3, 2, generator=g) - 0.5) * 0.5
    fvz = -(torch.rand(B, F, 3, generator=g) * 2 + 0.5)
    if behind:
        fvz[:, ::5] = fvz[:, ::5].abs()            # entirely behind the camera
        fvz[:, 1::7, 0] = fvz[:, 1::7, 0].abs()    # straddling z = 0
    fvi[:, 3::11, 2] = fvi[:, 3::11, 1]            # degenerate (two equal vertices)
    return fvz, fvi


@pytest.mark.parametrize("H,W,behind", [(24, 24, False), (17, 31, True), (40, 16, True)])
def test_traversals_agree(H, W, behind):
    fvz, fvi = _random_soup(2, 150, 7 + H, behind)
    valid = torch.rand(2, 150, generator=torch.Generator().manual_seed(1)) > 0.2
    ref = kal.rasterize_buffers(H, W, fvz, fvi, valid, impl="brute")
    for impl in ("bbox", "torch"):
        out = kal.rasterize_buffers(H, W, fvz, fvi, valid, impl=impl)
        for a, b, n in zip(ref, out, ("face_idx", "bary", "depth")):
            assert torch.equal(a, b), f"{impl}: {n} differs from brute force"
    assert (ref[0] >= 0).any() and (ref[0] < 0).any()


def test_tie_goes_to_lowest_face_index():
    tri = torch.tensor([[[-0.5, -0.5], [0.5, -0.5], [0.0, 0.6]]])
    fvi = tri[None].repeat(1, 3, 1, 1)
    fvz = torch.full((1, 3, 3), -2.0)
    fvz[0, 2] = -3.0
    idx, _, depth = kal.rasterize_buffers(16, 16, fvz, fvi, impl="brute")
    assert set(idx.unique().tolist()) == {-1, 0}
    assert torch.all(depth[idx == 0] == -2.0)
    idx2, _, _ = kal.rasterize_buffers(16, 16, fvz.flip(1), fvi, impl="bbox")   # nearer face now last
    assert set(idx2.unique().tolist()) == {-1, 1}


def test_single_triangle_known_answer():
    # screen-aligned right triangle covering the lower-left half of NDC [-1,1]^2 at depth -2
    fvi = torch.tensor([[[[-1.0, -1.0], [1.0, -1.0], [-1.0, 1.0]]]])
    fvz = torch.full((1, 1, 3), -2.0)
    feats = torch.tensor([[[[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]]]])
    out, idx = kal.rasterize(8, 8, fvz, fvi, feats)
    jj, ii = torch.meshgrid(torch.arange(8), torch.arange(8), indexing="ij")
    x = (2 * ii + 1 - 8) / 8.0
    y = (8 - 2 * jj - 1) / 8.0
    inside = (x + y) <= 0
    assert torch.equal(idx[0] >= 0, inside)
    assert_close(out[0, ..., 0][inside], ((x + 1) / 2)[inside], "u")
    assert_close(out[0, ..., 1][inside], ((y + 1) / 2)[inside], "v")


def test_behind_camera_faces_rejected():
    fvi = torch.tensor([[[[-1.0, -1.0], [1.0, -1.0], [0.0, 1.0]]]])
    idx, _, _ = kal.rasterize_buffers(8, 8, torch.full((1, 1, 3), 2.0), fvi)
    assert (idx == -1).all()


def test_texture_mapping_texel_centres_and_constant():
    T = 8
    tex = rnd((1, 3, T, T), 5)
    jj, ii = torch.meshgrid(torch.arange(T), torch.arange(T), indexing="ij")
    uv = torch.stack([(ii + 0.5) / T, 1 - (jj + 0.5) / T], dim=-1)[None].float()
    for mode in ("nearest", "bilinear"):
        out = kal.texture_mapping(uv, tex, mode=mode)
        assert_close(out[0].permute(2, 0, 1), tex[0], f"texel centres ({mode})")
    const = torch.full((1, 2, T, T), 0.75)
    out = kal.texture_mapping(torch.rand(1, 5, 5, 2), const, mode="bilinear")
    assert_close(out, torch.full_like(out, 0.75), "constant texture")


def test_bilinear_gradient_sums_to_upstream():
    tex = rnd((1, 2, 16, 16), 1).requires_grad_(True)
    uv = torch.rand(1, 9, 9, 2, generator=torch.Generator().manual_seed(3))
    g = rnd((1, 9, 9, 2), 2)
    kal.texture_mapping(uv, tex, mode="bilinear").backward(g)
    assert_close(tex.grad.sum(dim=(0, 2, 3)), g.sum(dim=(0, 1, 2)), "sum of texture gradient", rtol=1e-4, atol=1e-4)


def _run_mirror_latent_paint(gd):
    verts, faces, uv = scene(str(gd["shape"]), float(gd["scale"]), float(gd["dy"]))
    tex = torch.tensor(gd["texture"]).requires_grad_(True)
    r = renderer_ref.LatentPaintRendererRef(dim=tuple(int(d) for d in gd["dims"]), interpolation_mode=str(gd["mode"]))
    image, mask = r.render_single_view_texture(verts, faces, uv, tex, elev=float(gd["elev"]), azim=float(gd["azim"]),
                                               radius=float(gd["radius"]), look_at_height=float(gd["dy"]),
                                               white_background=bool(gd["white"]))
    image.backward(torch.tensor(gd["grad_image"]))
    return r, image, mask, tex.grad


@pytest.mark.parametrize("case", ["lp_blub_nearest", "lp_blub_bilinear_white"])
@pytest.mark.parametrize("impl", ["bbox", "torch"])
def test_mirror_reproduces_golden_latent_paint(case, impl, monkeypatch):
    monkeypatch.setattr(kal, "RASTER_IMPL", impl)
    gd = load_golden(case)
    r, image, mask, grad = _run_mirror_latent_paint(gd)
    assert np.array_equal(r.last["face_idx"].numpy(), gd["face_idx"])
    assert np.array_equal(mask.numpy(), gd["mask"])
    assert np.array_equal(image.detach().numpy(), gd["image"])
    assert np.array_equal(r.last["uv"].numpy(), gd["uv"])
    assert_close(grad, gd["grad_texture"], "grad_texture")


def test_mirror_reproduces_golden_env_sphere():
    gd = load_golden("lp_env_sphere_colors")
    verts, faces, _ = scene("env_sphere", 1.0, 0.0)
    import latent_nerf_test_b200 as lp
    m = lp.meshio.find_shape("env_sphere")
    colors = torch.tensor(gd["colors"]).requires_grad_(True)
    r = renderer_ref.LatentPaintRendererRef(dim=tuple(int(d) for d in gd["dims"]))
    image, mask = r.render_single_view(m.vertices, m.faces, colors, elev=float(gd["elev"]), azim=float(gd["azim"]),
                                       radius=float(gd["radius"]), look_at_height=0.25)
    image.backward(torch.tensor(gd["grad_image"]))
    assert np.array_equal(r.last["face_idx"].numpy(), gd["face_idx"])
    assert np.array_equal(image.detach().numpy(), gd["image"])
    assert_close(colors.grad, gd["grad_colors"], "grad_colors")
    assert mask.mean() == 1.0        # the camera sits inside the sphere


@pytest.mark.parametrize("case", ["mesh_sphere_body_b3", "mesh_teddy_head_white"])
def test_mirror_reproduces_golden_mesh(case):
    gd = load_golden(case)
    verts, faces, uv = scene(str(gd["shape"]), 1.0, 0.0)
    tex = torch.tensor(gd["texture"]).requires_grad_(True)
    dims = tuple(int(d) for d in gd["dims"])
    r = renderer_ref.LatentPaintMeshRendererRef(dim=dims, interpolation_mode="bilinear")
    radius = torch.tensor(gd["radius"]) if gd["radius"].ndim else float(gd["radius"])
    outs = r.render_single_view_texture(verts, faces, uv, tex, torch.tensor(gd["elev"]), torch.tensor(gd["azim"]), radius,
                                        dims=dims, white_background=bool(gd["white"]), is_body=bool(gd["is_body"]))
    outs[0].backward(torch.tensor(gd["grad_image"]))
    assert np.array_equal(r.last["face_idx"].numpy(), gd["face_idx"])
    for o, k in zip(outs, ("image", "mask", "normals", "lighting")):
        assert np.array_equal(o.detach().numpy(), gd[k]), k
    assert_close(tex.grad, gd["grad_texture"], "grad_texture")


def test_mesh_flavour_background_texel_leak():
    """Reference quirk: the mesh flavour does not mask the image, so every uncovered
    pixel samples — and back-propagates into — texel (row T-1, col 0) with weight 1."""
    verts, faces, uv = scene("sphere", 1.0, 0.0)
    T = 16
    tex = rnd((1, 2, T, T), 1).requires_grad_(True)
    r = renderer_ref.LatentPaintMeshRendererRef(dim=(32, 32))
    image, mask, _, _ = r.render_single_view_texture(verts, faces, uv, tex, torch.tensor([1.3]), torch.tensor([0.4]),
                                                     torch.tensor([2.2]), dims=(32, 32), is_body=True)
    image.backward(torch.ones_like(image))
    empty = int((r.last["face_idx"] < 0).sum())
    assert empty > 0
    assert (tex.grad[0, :, T - 1, 0] >= empty - 1e-3).all()
    assert_close(tex.grad.sum(dim=(0, 2, 3)), torch.full((2,), 32.0 * 32.0), "gradient mass", rtol=1e-4, atol=1e-2)


@pytest.mark.skipif(not reference_glue.available(), reason="needs /root/reference (build container only)")
def test_mirror_equals_reference_glue():
    """The mirror that travels to the GPU box vs the reference's real files, fresh inputs."""
    verts, faces, uv = scene("sphere", 0.6, 0.25)
    tex = rnd((1, 4, 32, 32), 11)
    R = reference_glue.load_latent_paint_renderer()
    a = R("cpu", dim=(40, 40), interpolation_mode="bilinear").render_single_view_texture(
        verts, faces, uv, tex, elev=0.9, azim=4.0, radius=1.3, look_at_height=0.25, white_background=True)
    b = renderer_ref.LatentPaintRendererRef(dim=(40, 40), interpolation_mode="bilinear").render_single_view_texture(
        verts, faces, uv, tex, elev=0.9, azim=4.0, radius=1.3, look_at_height=0.25, white_background=True)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    RM = reference_glue.load_latent_paint_mesh_renderer()
    verts, faces, uv = scene("sphere", 1.0, 0.0)
    th, ph, rad = torch.tensor([1.2, 1.7]), torch.tensor([0.3, 5.0]), torch.tensor([1.5, 2.2])
    a = RM("cpu", dim=(32, 32), interpolation_mode="bilinear").render_single_view_texture(
        verts, faces, uv, tex, th, ph, rad, dims=(32, 32), is_body=False)
    b = renderer_ref.LatentPaintMeshRendererRef(dim=(32, 32)).render_single_view_texture(
        verts, faces, uv, tex, th, ph, rad, dims=(32, 32), is_body=False)
    for x, y in zip(a, b):
        assert torch.equal(x, y)


def test_render_train_mirror_reproduces_golden():
    """Model-level composition (reference textured_mesh.py:187-220): the travelling mirror over the mirror renderer
    must reproduce the vectors frozen from the reference's real Renderer class."""
    gd = load_golden("lp_render_train_blub")
    verts, faces, uv = scene("blub", 0.6, 0.25)
    env_v, env_f, _ = scene_raw("env_sphere")
    tex = torch.tensor(gd["texture"]).requires_grad_(True)
    colors = torch.tensor(gd["colors"]).requires_grad_(True)
    ref = renderer_ref.LatentPaintRendererRef(dim=tuple(int(d) for d in gd["dims"]), interpolation_mode="bilinear")
    out = renderer_ref.render_train_ref(ref, verts, faces, uv, tex, env_v, env_f, colors, float(gd["elev"]),
                                        float(gd["azim"]), float(gd["radius"]), dy=0.25)
    out["image"].backward(torch.tensor(gd["grad_image"]))
    assert np.array_equal(out["mask"].numpy(), gd["mask"])
    for k in ("image", "background", "foreground"):
        assert_close(out[k], gd[k], k)
    assert_close(tex.grad, gd["grad_texture"], "grad_texture")
    assert_close(colors.grad, gd["grad_colors"], "grad_colors")
    # the composition really mixes both renders: background where the mask is 0, foreground where it is 1
    m = torch.tensor(gd["mask"]).bool().expand(1, 4, -1, -1)
    assert np.array_equal(gd["image"][m.numpy()], gd["foreground"][m.numpy()])
    assert np.array_equal(gd["image"][~m.numpy()], gd["background"][~m.numpy()])
    # resize branch: a 96 x 96 render comes back on the 64 x 64 latent grid (bicubic, :214-218)
    ref96 = renderer_ref.LatentPaintRendererRef(dim=(96, 96), interpolation_mode="bilinear")
    out96 = renderer_ref.render_train_ref(ref96, verts, faces, uv, tex.detach(), env_v, env_f, colors.detach(), 1.0, 0.7, 1.25)
    assert all(tuple(out96[k].shape[-2:]) == (64, 64) for k in ("image", "mask", "background", "foreground"))
