"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path, called through the
reference-shaped Renderer classes and through the C ABI, against the frozen reference-glue
vectors and against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): face_idx and the 0/1 mask bit-exact; pixels, float mask,
normals, lighting and gradients within rtol 1e-4 / atol 1e-5.
"""
import ctypes

import numpy as np
import pytest
import torch

import latent_nerf_test_b200 as lp
from latent_nerf_test_b200 import _lib, functional
from oracle import kaolin_shim as kal
from oracle import renderer_ref
from tests.common import (accumulation_terms, assert_close, assert_texture_grad_close, fp64_corner_texel, latent_paint_views, load_golden,
                          mesh_views, rnd, scene)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    kal.RASTER_IMPL = "bbox"


class _Mesh:
    def __init__(self, vertices, faces):
        self.vertices, self.faces = vertices, faces


# ------------------------------------------------------------------ golden vectors (reference glue)
@pytest.mark.parametrize("case", ["lp_blub_nearest", "lp_blub_bilinear_white"])
def test_golden_latent_paint_texture(case):
    gd = load_golden(case)
    verts, faces, uv = scene(str(gd["shape"]), float(gd["scale"]), float(gd["dy"]))
    tex = torch.tensor(gd["texture"], device=DEV).requires_grad_(True)
    r = lp.LatentPaintRenderer(DEV, dim=tuple(int(d) for d in gd["dims"]), interpolation_mode=str(gd["mode"]))
    r.keep_buffers = True
    image, mask = r.render_single_view_texture(verts.to(DEV), faces.to(DEV), uv.to(DEV), tex, elev=float(gd["elev"]),
                                               azim=float(gd["azim"]), radius=float(gd["radius"]),
                                               look_at_height=float(gd["dy"]), white_background=bool(gd["white"]))
    assert image.shape == gd["image"].shape and mask.shape == gd["mask"].shape and image.dtype == torch.float32
    assert np.array_equal(r.last_buffers["face_idx"].cpu().numpy(), gd["face_idx"])
    assert np.array_equal(mask.cpu().numpy(), gd["mask"])
    assert_close(image, gd["image"], "image")
    covered = torch.tensor(gd["face_idx"]) >= 0
    assert_close(r.last_buffers["uv"].cpu()[covered], torch.tensor(gd["uv"])[covered], "uv")
    image.backward(torch.tensor(gd["grad_image"], device=DEV))
    assert tex.grad.shape == gd["grad_texture"].shape
    assert_close(tex.grad, gd["grad_texture"], "grad_texture")


def test_golden_env_sphere_face_colours():
    gd = load_golden("lp_env_sphere_colors")
    m = lp.meshio.find_shape("env_sphere")
    colors = torch.tensor(gd["colors"], device=DEV).requires_grad_(True)
    r = lp.LatentPaintRenderer(DEV, dim=tuple(int(d) for d in gd["dims"]))
    r.keep_buffers = True
    image, mask = r.render_single_view(_Mesh(m.vertices.to(DEV), m.faces.to(DEV)), colors, elev=float(gd["elev"]),
                                       azim=float(gd["azim"]), radius=float(gd["radius"]), look_at_height=0.25)
    assert np.array_equal(r.last_buffers["face_idx"].cpu().numpy(), gd["face_idx"])
    assert np.array_equal(mask.cpu().numpy(), gd["mask"])
    assert_close(image, gd["image"], "image")
    image.backward(torch.tensor(gd["grad_image"], device=DEV))
    assert_close(colors.grad, gd["grad_colors"], "grad_colors")


@pytest.mark.parametrize("case", ["mesh_sphere_body_b3", "mesh_teddy_head_white"])
def test_golden_mesh_flavour(case):
    gd = load_golden(case)
    verts, faces, uv = scene(str(gd["shape"]), 1.0, 0.0)
    tex = torch.tensor(gd["texture"], device=DEV).requires_grad_(True)
    dims = tuple(int(d) for d in gd["dims"])
    r = lp.LatentPaintMeshRenderer(DEV, dim=dims, interpolation_mode="bilinear")
    r.keep_buffers = True
    radius = torch.tensor(gd["radius"]) if gd["radius"].ndim else float(gd["radius"])
    outs = r.render_single_view_texture(verts.to(DEV), faces.to(DEV), uv.to(DEV), tex, torch.tensor(gd["elev"]),
                                        torch.tensor(gd["azim"]), radius, dims=dims, white_background=bool(gd["white"]),
                                        is_body=bool(gd["is_body"]))
    assert np.array_equal(r.last_buffers["face_idx"].cpu().numpy(), gd["face_idx"])
    for o, k in zip(outs, ("image", "mask", "normals", "lighting")):
        assert o.shape == gd[k].shape, k
        assert_close(o, gd[k], k)
    assert np.array_equal((outs[1] > 0).cpu().numpy(), gd["face_idx"][:, None] >= 0)
    outs[0].backward(torch.tensor(gd["grad_image"], device=DEV))
    # unmasked flavour: texel (T-1, 0) sums every uncovered pixel -> derived accumulation bound around the fp64 sum
    bg = fp64_corner_texel(r.last_buffers["uv"], tex, gd["grad_image"], gd["face_idx"])
    # texels at the sphere's UV poles sum hundreds of pixels: per-texel accumulation term (summation order differs run to run)
    terms = accumulation_terms(r.last_buffers["uv"], gd["grad_image"], tex.shape)
    assert_texture_grad_close(tex.grad, gd["grad_texture"], "grad_texture", background=bg, terms=terms)


# ------------------------------------------------------------------ oracle, seeded random views
def _oracle_latent_paint(verts, faces, uv, tex, mode, dims, e, a, r, dy, white, grad):
    t = tex.detach().cpu().clone().requires_grad_(True)
    ref = renderer_ref.LatentPaintRendererRef(dim=dims, interpolation_mode=mode)
    image, mask = ref.render_single_view_texture(verts, faces, uv, t, elev=e, azim=a, radius=r, look_at_height=dy,
                                                 dims=dims, white_background=white)
    image.backward(grad.cpu())
    return image.detach(), mask, ref.last["face_idx"], t.grad


@pytest.mark.parametrize("shape,scale,dy,dims,T,C,mode,white", [
    ("blub", 0.6, 0.25, (64, 64), 128, 4, "nearest", False),       # config 1
    ("nascar", 0.6, 0.25, (512, 512), 1024, 3, "bilinear", False),  # config 2 geometry, one view at a time
    ("teddy", 0.6, 0.25, (200, 120), 64, 4, "bilinear", True),      # non-square, not a multiple of the tile
    ("sphere", 0.6, 0.25, (33, 47), 20, 5, "nearest", True),        # odd sizes, generic channel count, non-pow2 T
])
def test_latent_paint_vs_oracle(shape, scale, dy, dims, T, C, mode, white):
    verts, faces, uv = scene(shape, scale, dy)
    radius, theta, phi = latent_paint_views(3, seed=5)
    tex = rnd((1, C, T, T), 1, 0.4).to(DEV).requires_grad_(True)
    r = lp.LatentPaintRenderer(DEV, dim=dims, interpolation_mode=mode)
    r.keep_buffers = True
    vd, fd, ud = verts.to(DEV), faces.to(DEV), uv.to(DEV)
    for i in range(3):
        e, a, rad = float(theta[i]), float(phi[i]), float(radius[i])
        tex.grad = None
        image, mask = r.render_single_view_texture(vd, fd, ud, tex, elev=e, azim=a, radius=rad, look_at_height=dy,
                                                   dims=dims, white_background=white)
        assert image.shape == (1, C, dims[1], dims[0])
        g = rnd(tuple(image.shape), 20 + i)
        image.backward(g.to(DEV))
        oi, om, ofi, og = _oracle_latent_paint(verts, faces, uv, tex, mode, dims, e, a, rad, dy, white, g)
        assert torch.equal(r.last_buffers["face_idx"].cpu().long(), ofi), "face_idx must be bit-exact"
        assert torch.equal(mask.cpu(), om), "mask must be bit-exact"
        assert_close(image, oi, "image")
        assert_close(tex.grad, og, "grad_texture")
        assert int((ofi >= 0).sum()) > 0


def test_mesh_flavour_vs_oracle_batched():
    verts, faces, uv = scene("teddy", 1.0, 0.0)
    radius, theta, phi = mesh_views(4, seed=9)
    for is_body, dims, T in [(True, (64, 64), 512), (False, (96, 80), 128)]:
        tex = rnd((1, 4, T, T), 1, 0.4).to(DEV).requires_grad_(True)
        r = lp.LatentPaintMeshRenderer(DEV, dim=dims)
        r.keep_buffers = True
        outs = r.render_single_view_texture(verts.to(DEV), faces.to(DEV), uv.to(DEV), tex, theta, phi, radius, dims=dims,
                                            is_body=is_body)
        g = rnd(tuple(outs[0].shape), 3)
        outs[0].backward(g.to(DEV))
        t = tex.detach().cpu().clone().requires_grad_(True)
        ref = renderer_ref.LatentPaintMeshRendererRef(dim=dims)
        routs = ref.render_single_view_texture(verts, faces, uv, t, theta, phi, radius, dims=dims, is_body=is_body)
        routs[0].backward(g)
        assert torch.equal(r.last_buffers["face_idx"].cpu().long(), ref.last["face_idx"])
        for o, ro, k in zip(outs, routs, ("image", "mask", "normals", "lighting")):
            assert_close(o, ro, k)
        bg = fp64_corner_texel(ref.last["uv"], t, g, ref.last["face_idx"])
        assert_texture_grad_close(tex.grad, t.grad, "grad_texture", background=bg, terms=accumulation_terms(ref.last["uv"], g, t.shape))


def test_device_cameras_and_visibility_with_them():
    """Angle tensors on the GPU take lp_cameras_from_views: the matrices match the host math to
    fp32 rounding, and visibility stays bit-exact when the oracle is fed those same matrices."""
    verts, faces, uv = scene("sphere", 1.0, 0.0)
    radius, theta, phi = mesh_views(5, seed=2)
    r = lp.LatentPaintMeshRenderer(DEV, dim=(64, 64))
    r.keep_buffers = True
    cam_dev = r.get_camera_from_view(theta.to(DEV), phi.to(DEV), radius.to(DEV), -0.3)
    cam_host = renderer_ref._look_at_camera(theta, phi, radius, torch.tensor([-0.3]))
    assert_close(cam_dev, cam_host, "camera matrices", rtol=1e-5, atol=2e-6)
    tex = rnd((1, 4, 32, 32), 1).to(DEV)
    outs = r.render_single_view_texture(verts.to(DEV), faces.to(DEV), uv.to(DEV), tex, theta.to(DEV), phi.to(DEV),
                                        radius.to(DEV), is_body=True)
    M = r.last_buffers["camera"].cpu()
    fvc, fvi, fn = kal.prepare_vertices(verts, faces, kal.generate_perspective_projection(np.pi / 4), camera_transform=M)
    idx, _, _ = kal.rasterize_buffers(64, 64, fvc[..., -1], fvi, valid_faces=fn[..., -1].abs() > 0)
    assert torch.equal(r.last_buffers["face_idx"].cpu().long(), idx)
    assert outs[0].shape == (5, 4, 64, 64)


# ------------------------------------------------------------------ edge cases
def test_empty_view_and_huge_triangles():
    verts, faces, uv = scene("sphere", 0.6, 0.25)
    tex = rnd((1, 4, 16, 16), 1).to(DEV).requires_grad_(True)
    r = lp.LatentPaintRenderer(DEV, dim=(40, 40), interpolation_mode="bilinear")
    r.keep_buffers = True
    # the mesh is far outside the frustum: nothing is visible (checked with the oracle: 0 covered pixels)
    far = (verts + torch.tensor([0.0, -100.0, 0.0])).to(DEV)
    image, mask = r.render_single_view_texture(far, faces.to(DEV), uv.to(DEV), tex, elev=1.0, azim=0.5, radius=1.2)
    assert float(mask.sum()) == 0 and float(image.abs().sum()) == 0
    image.sum().backward()
    assert float(tex.grad.abs().sum()) == 0
    image, mask = r.render_single_view_texture(far, faces.to(DEV), uv.to(DEV), tex, elev=1.0, azim=0.5, radius=1.2,
                                               white_background=True)
    assert float((image - 1).abs().sum()) == 0
    # no near-plane clipping (kaolin semantics): faces straddling the camera plane smear over the frame
    image, mask = r.render_single_view_texture(verts.to(DEV), faces.to(DEV), uv.to(DEV), tex, elev=1.0, azim=0.5, radius=1.2,
                                               look_at_height=50.0)
    oi, om, ofi, _ = _oracle_latent_paint(verts, faces, uv, tex, "bilinear", (40, 40), 1.0, 0.5, 1.2, 50.0, False,
                                          torch.zeros(1, 4, 40, 40))
    assert torch.equal(r.last_buffers["face_idx"].cpu().long(), ofi) and float(om.sum()) == 1600
    # a few triangles filling the whole frame exercise the coarse pyramid levels
    big_v = torch.tensor([[-9.0, -9.0, 0.0], [9.0, -9.0, 0.0], [0.0, 9.0, 0.0], [-9.0, -9.0, 0.3], [9.0, -9.0, 0.3],
                          [-9.0, 9.0, 0.3]])
    big_f = torch.tensor([[0, 1, 2], [3, 4, 5]])
    big_uv = torch.rand(1, 2, 3, 2, generator=torch.Generator().manual_seed(0))
    for dims in [(40, 40), (300, 200)]:
        image, mask = r.render_single_view_texture(big_v.to(DEV), big_f.to(DEV), big_uv.to(DEV), tex, elev=1.4, azim=0.1,
                                                   radius=2.0, dims=dims)
        oi, om, ofi, _ = _oracle_latent_paint(big_v, big_f, big_uv, tex, "bilinear", dims, 1.4, 0.1, 2.0, 0.0, False,
                                              torch.zeros(1, 4, dims[1], dims[0]))
        assert torch.equal(r.last_buffers["face_idx"].cpu().long(), ofi)
        assert_close(image, oi, "image")
        assert float(om.mean()) > 0.3 and set(ofi.unique().tolist()) >= {0, 1}


def test_degenerate_and_behind_camera_faces():
    """env_sphere encloses the camera: thousands of faces behind / straddling the image plane."""
    m = lp.meshio.find_shape("env_sphere")
    colors = rnd((1, m.faces.shape[0], 3, 4), 3)
    for flag in (True, False):
        r = lp.LatentPaintRenderer(DEV, dim=(48, 48))
        r.keep_buffers, r.reject_behind_camera = True, flag
        image, mask = r.render_single_view(_Mesh(m.vertices.to(DEV), m.faces.to(DEV)), colors.to(DEV), elev=0.8, azim=2.0,
                                           radius=1.4, look_at_height=0.25)
        kal.REJECT_BEHIND_CAMERA = flag
        try:
            ref = renderer_ref.LatentPaintRendererRef(dim=(48, 48))
            oi, om = ref.render_single_view(m.vertices, m.faces, colors, elev=0.8, azim=2.0, radius=1.4, look_at_height=0.25)
        finally:
            kal.REJECT_BEHIND_CAMERA = True
        assert torch.equal(r.last_buffers["face_idx"].cpu().long(), ref.last["face_idx"]), f"reject_behind={flag}"
        assert_close(image, oi, "image")


def test_high_poly_stress_against_oracle():
    """config-4 style: subdivided sphere (81 920 faces, sub-pixel triangles) — bins hold hundreds of faces."""
    verts, faces, uv = scene("sphere", 0.6, 0.25, subdivide=3)
    tex = rnd((1, 3, 256, 256), 1).to(DEV).requires_grad_(True)
    r = lp.LatentPaintRenderer(DEV, dim=(256, 256), interpolation_mode="bilinear")
    r.keep_buffers = True
    image, mask = r.render_single_view_texture(verts.to(DEV), faces.to(DEV), uv.to(DEV), tex, elev=1.1, azim=0.3, radius=1.2,
                                               look_at_height=0.25)
    g = rnd(tuple(image.shape), 2)
    image.backward(g.to(DEV))
    oi, om, ofi, og = _oracle_latent_paint(verts, faces, uv, tex, "bilinear", (256, 256), 1.1, 0.3, 1.2, 0.25, False, g)
    assert torch.equal(r.last_buffers["face_idx"].cpu().long(), ofi)
    assert_close(image, oi, "image")
    assert_close(tex.grad, og, "grad_texture")


# ------------------------------------------------------------------ size-independent properties at config-2 size
def test_config2_properties():
    verts, faces, uv = scene("nascar", 0.6, 0.25)
    vd, fd, ud = verts.to(DEV), faces.to(DEV), uv.to(DEV)
    r = lp.LatentPaintRenderer(DEV, dim=(512, 512), interpolation_mode="bilinear")
    r.keep_buffers = True
    radius, theta, phi = latent_paint_views(8, seed=0)
    const = torch.full((1, 3, 1024, 1024), 0.625, device=DEV)
    tex = rnd((1, 3, 1024, 1024), 1, 0.4).to(DEV).requires_grad_(True)
    for i in range(8):
        e, a, rad = float(theta[i]), float(phi[i]), float(radius[i])
        image, mask = r.render_single_view_texture(vd, fd, ud, const, elev=e, azim=a, radius=rad, look_at_height=0.25)
        fi = r.last_buffers["face_idx"].clone()
        assert_close(image, 0.625 * mask.expand_as(image), "constant texture → constant · mask")
        assert torch.equal(mask[0, 0] > 0, fi[0] >= 0)
        depth = r.last_buffers["depth"]
        assert bool((depth[fi >= 0] < 0).all()) and bool((depth[fi < 0] == 0).all())
        bary = r.last_buffers["bary"]
        assert_close(bary.sum(-1)[fi >= 0], torch.ones(int((fi >= 0).sum())), "barycentrics sum to one", atol=1e-5)
        # determinism: same buffers bit for bit on a second run
        image2, _ = r.render_single_view_texture(vd, fd, ud, const, elev=e, azim=a, radius=rad, look_at_height=0.25)
        assert torch.equal(image, image2) and torch.equal(fi, r.last_buffers["face_idx"])
        # bilinear taps sum to one ⇒ the gradient mass equals the masked upstream mass
        tex.grad = None
        img, msk = r.render_single_view_texture(vd, fd, ud, tex, elev=e, azim=a, radius=rad, look_at_height=0.25)
        g = rnd(tuple(img.shape), 40 + i).to(DEV)
        img.backward(g)
        assert_close(tex.grad.sum(dim=(0, 2, 3)), (g * msk).sum(dim=(0, 2, 3)), "gradient mass", rtol=1e-3, atol=1e-2)
        # linearity of the backward in the upstream gradient
        g1 = tex.grad.clone()
        tex.grad = None
        img, _ = r.render_single_view_texture(vd, fd, ud, tex, elev=e, azim=a, radius=rad, look_at_height=0.25)
        img.backward(2.0 * g)
        assert_close(tex.grad, 2.0 * g1, "backward is linear", rtol=1e-4, atol=1e-4)


# ------------------------------------------------------------------ C ABI directly
def test_c_abi_error_codes_on_device():
    L = _lib.lib()
    verts, faces, uv = scene("sphere", 0.6, 0.25)
    v = verts.to(DEV).contiguous()
    f = faces.to(DEV, torch.int32).contiguous()
    u = uv.to(DEV).reshape(-1, 3, 2).contiguous()
    cam = lp.camera.camera_from_view(torch.tensor(1.0), torch.tensor(0.5), 1.3, 0.25).to(DEV).contiguous()
    tex = rnd((1, 4, 16, 16), 1).to(DEV)
    image = torch.empty(1, 4, 32, 32, device=DEV)
    mask = torch.empty(1, 1, 32, 32, device=DEV)
    a = _lib.LpForwardArgs()
    a.verts, a.faces, a.V, a.F = v.data_ptr(), f.data_ptr(), v.shape[0], f.shape[0]
    a.cameras, a.B, a.H, a.W = cam.data_ptr(), 1, 32, 32
    a.proj[0] = a.proj[1] = 1.7320508; a.proj[2] = -1.0
    a.multiplier, a.eps, a.flags = 1000.0, 1e-8, _lib.LP_FLAG_MASK_IMAGE | _lib.LP_FLAG_REJECT_BEHIND
    a.face_uv, a.texture, a.C, a.Th, a.Tw, a.interp = u.data_ptr(), tex.data_ptr(), 4, 16, 16, 0
    a.image, a.mask = image.data_ptr(), mask.data_ptr()
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert L.lp_render_forward(ctypes.byref(a), stream) == _lib.LP_ERR_WORKSPACE
    need = L.lp_workspace_bytes(1, f.shape[0], 32, 32)
    ws = torch.empty(need - 1, dtype=torch.uint8, device=DEV)
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    assert L.lp_render_forward(ctypes.byref(a), stream) == _lib.LP_ERR_WORKSPACE
    ws = torch.empty(need, dtype=torch.uint8, device=DEV)
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    a.interp = 3                                # 0 / 1 / 2 = nearest / bilinear / bicubic
    assert L.lp_render_forward(ctypes.byref(a), stream) == _lib.LP_ERR_UNSUPPORTED
    a.interp = 0
    assert L.lp_render_forward(ctypes.byref(a), stream) == _lib.LP_OK
    assert L.lp_last_launch_count() == 4        # setup + binning, large-face binning, footprint classification, footprint kernel
    torch.cuda.synchronize()
    assert float(mask.sum()) > 0
    with pytest.raises(AssertionError):
        lp.LatentPaintRenderer(DEV, dim=(32, 32), interpolation_mode="lanczos")
    with pytest.raises(RuntimeError, match="no CPU path"):
        lp.LatentPaintRenderer(DEV, dim=(32, 32)).render_single_view_texture(v, faces.to(DEV), uv.to(DEV), tex.cpu())


def test_host_buffer_step_matches_device_path():
    """lp_render_step_host (what bench.py's e2e times) gives the same image / gradient as the
    autograd path."""
    L = _lib.lib()
    verts, faces, uv = scene("nascar", 0.6, 0.25)
    B, H, W, C, T = 2, 128, 128, 3, 256
    radius, theta, phi = latent_paint_views(B, seed=3)
    cams = torch.cat([lp.camera.camera_from_view(theta[i], phi[i], float(radius[i]), 0.25) for i in range(B)]).contiguous()
    tex = rnd((1, C, T, T), 1, 0.4).to(DEV).requires_grad_(True)
    g = rnd((B, C, H, W), 2)
    r = lp.LatentPaintRenderer(DEV, dim=(W, H), interpolation_mode="bilinear")
    imgs = []
    for i in range(B):
        img, _ = r.render_single_view_texture(verts.to(DEV), faces.to(DEV), uv.to(DEV), tex, elev=float(theta[i]),
                                              azim=float(phi[i]), radius=float(radius[i]), look_at_height=0.25)
        img.backward(g[i:i + 1].to(DEV))
        imgs.append(img.detach())
    from bench import HostStep
    hs = HostStep(verts, faces, uv, tex.detach(), B, H, W, "bilinear", np.pi / 3)
    image_h, mask_h, grad_h = hs.step(cams, g)
    assert_close(image_h, torch.cat(imgs), "image through host buffers")
    # two different summation orders of the same per-view contributions (views one by one through autograd vs one
    # batched scatter): each texel sums a handful of fp32 terms, so the pixel tolerance applies unchanged
    assert_close(grad_h, tex.grad[0], "grad_texture through host buffers")


# ------------------------------------------------------------------ kaolin-namespaced operator API
def test_kaolin_compat_runs_the_reference_glue_on_the_kernels():
    """The reference's glue (its travelling mirror, proven equal to the real files on CPU) executed over
    ``latent_nerf_test_b200.kaolin_compat`` on the GPU vs the same glue over the CPU oracle."""
    kc = lp.kaolin_compat.make_module()
    verts, faces, uv = scene("blub", 0.6, 0.25)
    for mode, white, dims in [("nearest", False, (64, 64)), ("bilinear", True, (96, 72))]:
        tex = rnd((1, 4, 128, 128), 1, 0.4)
        g = rnd((1, 4, dims[1], dims[0]), 2)
        view = dict(elev=1.0, azim=0.7, radius=1.25, look_at_height=0.25, dims=dims, white_background=white)
        tg = tex.to(DEV).requires_grad_(True)
        rg = renderer_ref.LatentPaintRendererRef(dim=dims, interpolation_mode=mode, kal=kc, device=DEV)
        ig, mg = rg.render_single_view_texture(verts.to(DEV), faces.to(DEV), uv.to(DEV), tg, **view)
        ig.backward(g.to(DEV))
        tc = tex.clone().requires_grad_(True)
        rc = renderer_ref.LatentPaintRendererRef(dim=dims, interpolation_mode=mode)
        ic, mc = rc.render_single_view_texture(verts, faces, uv, tc, **view)
        ic.backward(g)
        assert torch.equal(rg.last["face_idx"].cpu(), rc.last["face_idx"])
        assert torch.equal(mg.cpu(), mc)
        assert_close(ig, ic, f"image ({mode})")
        assert_close(tg.grad, tc.grad, f"grad_texture ({mode})")
    # per-face-vertex colours: rasterize backward into the face features
    m = lp.meshio.find_shape("env_sphere")
    colors = rnd((1, m.faces.shape[0], 3, 4), 3)
    g = rnd((1, 4, 48, 48), 4)
    cg = colors.to(DEV).requires_grad_(True)
    rg = renderer_ref.LatentPaintRendererRef(dim=(48, 48), kal=kc, device=DEV)
    ig, _ = rg.render_single_view(m.vertices.to(DEV), m.faces.to(DEV), cg, elev=0.8, azim=2.0, radius=1.4, look_at_height=0.25)
    ig.backward(g.to(DEV))
    cc = colors.clone().requires_grad_(True)
    rc = renderer_ref.LatentPaintRendererRef(dim=(48, 48))
    ic, _ = rc.render_single_view(m.vertices, m.faces, cc, elev=0.8, azim=2.0, radius=1.4, look_at_height=0.25)
    ic.backward(g)
    assert torch.equal(rg.last["face_idx"].cpu(), rc.last["face_idx"])
    assert_close(ig, ic, "face-colour image")
    assert_close(cg.grad, cc.grad, "grad_colors")
    # mesh flavour: dibr_rasterization with a feature list, per-view texture copies, SH lighting
    verts, faces, uv = scene("teddy", 1.0, 0.0)
    radius, theta, phi = mesh_views(3, seed=6)
    tex = rnd((1, 4, 64, 64), 1, 0.4)
    g = rnd((3, 4, 64, 64), 2)
    tg = tex.to(DEV).requires_grad_(True)
    rg = renderer_ref.LatentPaintMeshRendererRef(dim=(64, 64), kal=kc, device=DEV)
    og = rg.render_single_view_texture(verts.to(DEV), faces.to(DEV), uv.to(DEV), tg, theta, phi, radius, dims=(64, 64))
    og[0].backward(g.to(DEV))
    tc = tex.clone().requires_grad_(True)
    rc = renderer_ref.LatentPaintMeshRendererRef(dim=(64, 64))
    oc = rc.render_single_view_texture(verts, faces, uv, tc, theta, phi, radius, dims=(64, 64))
    oc[0].backward(g)
    assert torch.equal(rg.last["face_idx"].cpu(), rc.last["face_idx"])
    for a, b, k in zip(og, oc, ("image", "mask", "normals", "lighting")):
        assert_close(a, b, k)
    bg = fp64_corner_texel(rc.last["uv"], tc, g, rc.last["face_idx"])
    assert_texture_grad_close(tg.grad, tc.grad, "grad_texture (mesh flavour)", background=bg, terms=accumulation_terms(rc.last["uv"], g, tc.shape))


def test_split_forward_equals_fused_forward():
    """lp_render_prepare + lp_render_raster + lp_render_shade (the pipelined form bench.py overlaps across
    steps) must give exactly the buffers of the fused lp_render_forward."""
    from bench import DeviceStep, WORKLOADS, cameras_for, make_views
    L = _lib.lib()
    # W % 4 == 0 takes the four-pixels-per-thread texture-fetch kernel, W = 202 the one-pixel one
    for flavour_white, interp, width in ((False, "bilinear", 208), (True, "bilinear", 208), (False, "nearest", 208),
                                         (True, "bilinear", 202), (False, "nearest", 202)):
        w = dict(WORKLOADS["c2"], B=2, H=160, W=width, T=256, interp=interp)
        verts, faces, uv = scene(w["shape"], w["scale"], w["dy"])
        geom = (verts.to(DEV).float().contiguous(), faces.to(DEV, torch.int32).contiguous(),
                uv.to(DEV).float().reshape(-1, 3, 2).contiguous())
        radius, theta, phi = make_views(w["B"], 5)
        cams = cameras_for(radius, theta, phi, w["dy"])
        a, b_ = DeviceStep(geom, w, cams, 1, torch.device(DEV)), DeviceStep(geom, w, cams, 1, torch.device(DEV))
        if flavour_white:
            for st in (a, b_):
                st.fwd.flags |= _lib.LP_FLAG_WHITE_BACKGROUND
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(L.lp_render_forward(ctypes.byref(a.fwd), stream))
        b_.image.fill_(-7.0)
        for fn in (L.lp_render_prepare, L.lp_render_raster, L.lp_render_shade):
            _lib.check(fn(ctypes.byref(b_.fwd), stream))
        torch.cuda.synchronize()
        assert torch.equal(a.image, b_.image) and torch.equal(a.mask, b_.mask) and torch.equal(a.footprint_any, b_.footprint_any)
        live = a.footprint_any.bool().repeat_interleave(4, 1).repeat_interleave(8, 2)[:, :w["H"], :w["W"]]
        ua, ub = a.uv[live], b_.uv[live]                        # uncovered pixels of live tiles carry the NaN marker
        assert torch.equal(torch.isnan(ua), torch.isnan(ub)) and torch.equal(torch.nan_to_num(ua), torch.nan_to_num(ub))
        assert float(a.mask.sum()) > 0


def test_random_scenes_visibility_bit_exact():
    """Seeded sweep over meshes, odd resolutions, camera distances (including cameras close enough that faces
    straddle the image plane) and both flavours: the visibility buffer must equal the oracle's bit for bit."""
    g = torch.Generator().manual_seed(1234)
    meshes = {name: scene(name, 0.6, 0.25) for name in ("sphere", "blub", "teddy", "nascar")}
    tex = rnd((1, 3, 37, 53), 1).to(DEV)                     # non-square, non-power-of-two texture
    checked = 0
    for it in range(24):
        name = list(meshes)[it % len(meshes)]
        verts, faces, uv = meshes[name]
        W = int(torch.randint(8, 200, (1,), generator=g)); H = int(torch.randint(8, 200, (1,), generator=g))
        elev = float(torch.rand(1, generator=g) * 2.6 + 0.25); azim = float(torch.rand(1, generator=g) * 6.28)
        radius = float(torch.rand(1, generator=g) * 1.4 + (0.45 if it % 6 == 5 else 0.9))   # 0.45: inside the mesh's bounding sphere
        mode = "bilinear" if it % 2 else "nearest"
        r = lp.LatentPaintRenderer(DEV, dim=(W, H), interpolation_mode=mode)
        r.keep_buffers = True
        image, mask = r.render_single_view_texture(verts.to(DEV), faces.to(DEV), uv.to(DEV), tex, elev=elev, azim=azim,
                                                   radius=radius, look_at_height=0.25, white_background=bool(it % 3 == 0))
        ref = renderer_ref.LatentPaintRendererRef(dim=(W, H), interpolation_mode=mode)
        oi, om = ref.render_single_view_texture(verts, faces, uv, tex.cpu(), elev=elev, azim=azim, radius=radius,
                                                look_at_height=0.25, white_background=bool(it % 3 == 0))
        assert torch.equal(r.last_buffers["face_idx"].cpu().long(), ref.last["face_idx"]), (it, name, W, H, radius)
        assert torch.equal(mask.cpu(), om), (it, name)
        if mode == "bilinear":                                # nearest on a non-power-of-two texture may flip exact ties
            assert_close(image, oi, f"image {it}")
        checked += int((ref.last["face_idx"] >= 0).sum())
    assert checked > 10000


def test_allreduce_unpack_single_rank_is_the_unpack():
    """lp_allreduce_unpack with world = 1 over plain peer pointers: the planar gradient must equal the
    texel-interleaved accumulation buffer transposed (the exchange kernel's unpack half), and a backward with
    LP_FLAG_GRAD_INTERLEAVED followed by it must equal the ordinary backward."""
    from bench import DeviceStep, WORKLOADS, cameras_for, make_views
    L = _lib.lib()
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for C, T in ((3, 64), (4, 32), (1, 16)):
        ntex = T * T
        buf = torch.zeros(C * ntex + 4 * ntex, device=DEV)                 # planar gradient, then the accumulation buffer
        acc = torch.randn(ntex, 4, device=DEV, generator=torch.Generator(device=DEV).manual_seed(C))
        buf[C * ntex:].copy_(acc.reshape(-1))
        ptrs = torch.tensor([buf.data_ptr()], dtype=torch.int64, device=DEV)
        _lib.check(L.lp_allreduce_unpack(None, ctypes.c_void_p(ptrs.data_ptr()), 4 * C * ntex, 0, ntex, C, 0, 1, stream))
        torch.cuda.synchronize()
        assert torch.equal(buf[:C * ntex].view(C, ntex), acc[:, :C].t().contiguous())
    assert L.lp_allreduce_unpack(None, ctypes.c_void_p(ptrs.data_ptr()), 0, 0, 6, 3, 0, 1, stream) == _lib.LP_ERR_BAD_ARG

    w = dict(WORKLOADS["c2"], B=2, H=96, W=96, T=128)
    verts, faces, uv = scene(w["shape"], w["scale"], w["dy"])
    geom = (verts.to(DEV).float().contiguous(), faces.to(DEV, torch.int32).contiguous(), uv.to(DEV).float().reshape(-1, 3, 2).contiguous())
    radius, theta, phi = make_views(w["B"], 9)
    cams = cameras_for(radius, theta, phi, w["dy"])
    ref = DeviceStep(geom, w, cams, 1, torch.device(DEV))
    ntex, C = w["T"] ** 2, w["C"]
    both = torch.zeros(C * ntex + 4 * ntex, device=DEV)
    fused = DeviceStep(geom, w, cams, 1, torch.device(DEV), grad_tex=both[:C * ntex].view(C, w["T"], w["T"]), accum=both[C * ntex:])
    with torch.cuda.stream(torch.cuda.current_stream()):
        ref.run(); fused.run()
    ptrs = torch.tensor([both.data_ptr()], dtype=torch.int64, device=DEV)
    _lib.check(L.lp_allreduce_unpack(None, ctypes.c_void_p(ptrs.data_ptr()), 4 * C * ntex, 0, ntex, C, 0, 1, stream))
    torch.cuda.synchronize()
    assert float(ref.grad_tex.abs().sum()) > 0
    assert_close(fused.grad_tex, ref.grad_tex, "gradient through the fused exchange path", rtol=1e-5, atol=1e-6)


def test_exchange_step_single_rank_bulk_copies_and_register_loads():
    """lp_exchange_step with world = 1 over a plain pointer (the in-kernel handshakes have no peer to wait for): the peer
    form — bulk asynchronous copies (cp.async.bulk + mbarrier) into shared memory — and the register-load form must both
    leave the planar gradient = the texel-interleaved accumulation buffer transposed, over repeated launches with
    changing grids (the epoch lives in the flag block, not in per-CTA state), and the optimiser epilogue must equal
    torch.optim.Adam on that gradient."""
    L = _lib.lib()
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    flag_floats = _lib.LP_EXCHANGE_FLAG_BYTES // 4
    try:
        for bulk in (1, 0):
            _lib.check(L.lp_set_option(_lib.LP_OPT_EXCHANGE_BULK, bulk))
            for C, T in ((3, 256), (4, 100), (1, 32)):
                ntex = T * T
                # planar gradient | accumulation buffer | planar parameters | flag block
                buf = torch.zeros(C * ntex + 4 * ntex + C * ntex + flag_floats, device=DEV)
                ptrs = torch.tensor([buf.data_ptr()], dtype=torch.int64, device=DEV)
                a = _lib.LpExchangeArgs()
                a.multicast_base, a.buffer_ptrs_dev = None, ctypes.c_void_p(ptrs.data_ptr())
                a.accum_offset, a.grad_offset = 4 * C * ntex, 0
                a.param_offset, a.flags_offset = 4 * (C * ntex + 4 * ntex), 4 * (2 * C * ntex + 4 * ntex)
                a.ntex, a.C, a.rank, a.world = ntex, C, 0, 1
                gen = torch.Generator(device=DEV).manual_seed(10 * C + bulk)
                for it, ctas in enumerate((0, 7, 300, 0)):
                    _lib.check(L.lp_set_option(_lib.LP_OPT_EXCHANGE_CTAS, ctas))
                    acc = torch.randn(ntex, 4, device=DEV, generator=gen)
                    buf[C * ntex:C * ntex + 4 * ntex].copy_(acc.reshape(-1))
                    _lib.check(L.lp_exchange_step(ctypes.byref(a), stream))
                    torch.cuda.synchronize()
                    assert torch.equal(buf[:C * ntex].view(C, ntex), acc[:, :C].t().contiguous()), (bulk, C, T, it)
                _lib.check(L.lp_set_option(_lib.LP_OPT_EXCHANGE_CTAS, 0))
                # optimiser epilogue: parameters in the allocation, state local
                p0 = 0.4 * torch.randn(C, ntex, device=DEV, generator=gen)
                buf[C * ntex + 4 * ntex:2 * C * ntex + 4 * ntex].copy_(p0.reshape(-1))
                m, v = torch.zeros(C * ntex, device=DEV), torch.zeros(C * ntex, device=DEV)
                pref = torch.nn.Parameter(p0.clone())
                opt = torch.optim.Adam([pref], lr=0.01, betas=(0.9, 0.99), eps=1e-15)
                a.adam, a.exp_avg, a.exp_avg_sq = 1, m.data_ptr(), v.data_ptr()
                a.lr, a.beta1, a.beta2, a.eps = 0.01, 0.9, 0.99, 1e-15
                for step in (1, 2, 3):
                    acc = torch.randn(ntex, 4, device=DEV, generator=gen)
                    buf[C * ntex:C * ntex + 4 * ntex].copy_(acc.reshape(-1))
                    a.step = step
                    _lib.check(L.lp_exchange_step(ctypes.byref(a), stream))
                    pref.grad = acc[:, :C].t().contiguous()
                    opt.step()
                    torch.cuda.synchronize()
                    assert_close(buf[C * ntex + 4 * ntex:2 * C * ntex + 4 * ntex].view(C, ntex), pref.detach(),
                                 f"sharded Adam epilogue (bulk={bulk}, C={C})", rtol=1e-5, atol=1e-6)
    finally:
        L.lp_set_option(_lib.LP_OPT_EXCHANGE_BULK, 1)
        L.lp_set_option(_lib.LP_OPT_EXCHANGE_CTAS, 0)


def test_fused_render_train_composition():
    """SURVEY.md §8 f rank 1: object render + environment-sphere render + pred_back * (1 - mask) + pred_features * mask
    (reference textured_mesh.py:187-220) through the fused composition, against the vectors frozen from the
    reference's real Renderer class, against the unfused calls on the GPU, and the resize branch against the oracle."""
    gd = load_golden("lp_render_train_blub")
    verts, faces, uv = scene("blub", 0.6, 0.25)
    env = lp.meshio.find_shape("env_sphere")
    envm = _Mesh(env.vertices.to(DEV), env.faces.to(DEV))
    objm = _Mesh(verts.to(DEV), faces.to(DEV))
    view = dict(theta=float(gd["elev"]), phi=float(gd["azim"]), radius=float(gd["radius"]))
    tex = torch.tensor(gd["texture"], device=DEV).requires_grad_(True)
    colors = torch.tensor(gd["colors"], device=DEV).requires_grad_(True)
    r = lp.LatentPaintRenderer(DEV, dim=tuple(int(d) for d in gd["dims"]), interpolation_mode="bilinear")
    out = lp.textured_mesh.render_train(r, objm, uv.to(DEV), tex, envm, colors, dy=0.25, **view)
    assert np.array_equal(out["mask"].cpu().numpy(), gd["mask"])
    for k in ("image", "background", "foreground"):
        assert_close(out[k], gd[k], k)
    out["image"].backward(torch.tensor(gd["grad_image"], device=DEV))
    assert_close(tex.grad, gd["grad_texture"], "grad_texture")
    assert_close(colors.grad, gd["grad_colors"], "grad_colors")

    # the unfused sequence on the same kernels gives the same bits forward (same expression, same order)
    fg, mask = r.render_single_view_texture(objm.vertices, objm.faces, uv.to(DEV), tex.detach(), elev=view["theta"],
                                            azim=view["phi"], radius=view["radius"], look_at_height=0.25)
    bg, _ = r.render_single_view(envm, colors.detach(), elev=view["theta"], azim=view["phi"], radius=view["radius"], look_at_height=0.25)
    assert torch.equal(out["foreground"], fg) and torch.equal(out["background"], bg)
    assert torch.equal(out["image"], bg * (1 - mask) + fg * mask)

    # gradients arriving on the other outputs as well (not the reference's loss, but autograd allows it)
    tex2 = tex.detach().clone().requires_grad_(True); col2 = colors.detach().clone().requires_grad_(True)
    o2 = lp.textured_mesh.render_train(r, objm, uv.to(DEV), tex2, envm, col2, dy=0.25, **view)
    g1, g2, g3 = (rnd((1, 4, 64, 64), s).to(DEV) for s in (11, 12, 13))
    (o2["image"] * g1 + o2["background"] * g2 + o2["foreground"] * g3).sum().backward()
    tex3 = tex.detach().clone().requires_grad_(True); col3 = colors.detach().clone().requires_grad_(True)
    fg3, m3 = r.render_single_view_texture(objm.vertices, objm.faces, uv.to(DEV), tex3, elev=view["theta"], azim=view["phi"],
                                           radius=view["radius"], look_at_height=0.25)
    bg3, _ = r.render_single_view(envm, col3, elev=view["theta"], azim=view["phi"], radius=view["radius"], look_at_height=0.25)
    ((bg3 * (1 - m3) + fg3 * m3) * g1 + bg3 * g2 + fg3 * g3).sum().backward()
    assert_close(tex2.grad, tex3.grad, "grad_texture, all outputs used")
    assert_close(col2.grad, col3.grad, "grad_colors, all outputs used")

    # resize branch (render grid != 64 in latent mode): bicubic to 64 x 64, against the CPU oracle
    r96 = lp.LatentPaintRenderer(DEV, dim=(96, 96), interpolation_mode="bilinear")
    o96 = lp.textured_mesh.render_train(r96, objm, uv.to(DEV), tex.detach(), envm, colors.detach(), dy=0.25, **view)
    ref96 = renderer_ref.LatentPaintRendererRef(dim=(96, 96), interpolation_mode="bilinear")
    e96 = renderer_ref.render_train_ref(ref96, verts, faces, uv, tex.detach().cpu(), env.vertices, env.faces,
                                        colors.detach().cpu(), view["theta"], view["phi"], view["radius"], dy=0.25)
    for k in ("image", "mask", "background", "foreground"):
        assert tuple(o96[k].shape[-2:]) == (64, 64)
        # the bicubic filter has negative lobes (weights in [-0.07, 0.6], |w| summing to <= 1.5625^2): it amplifies
        # the per-pixel fp32 differences of its 16 inputs by at most that factor
        assert_close(o96[k], e96[k], k + " (resized)", rtol=1e-4, atol=1e-5 * 1.5625 ** 2)


def test_fused_adam_matches_torch_adam():
    """SURVEY.md §8 f rank 4: the reference's optimiser (Adam, betas (0.9, 0.99), eps 1e-15, trainer.py:93-95) as one
    kernel, against torch.optim.Adam on the CPU (fp32 tolerance), planar gradients and the interleaved accumulation
    buffer of the vector-RED backward (unpack fused into the update)."""
    g = torch.Generator().manual_seed(5)
    for shape in ((1, 4, 32, 32), (1, 3, 17, 19)):            # 17 * 19 texels: the non-vector tail
        p_ref = torch.nn.Parameter(0.4 * torch.randn(*shape, generator=g))
        p_gpu = p_ref.detach().clone().to(DEV).requires_grad_(True)
        p_acc = p_ref.detach().clone().to(DEV)
        ref = torch.optim.Adam([p_ref], lr=0.01, betas=(0.9, 0.99), eps=1e-15)
        opt = lp.optim.FusedAdam([p_gpu], lr=0.01, betas=(0.9, 0.99), eps=1e-15)
        opt_acc = lp.optim.FusedAdam([p_acc], lr=0.01, betas=(0.9, 0.99), eps=1e-15)
        C, ntex = shape[1], shape[2] * shape[3]
        for step in range(6):
            grad = torch.randn(*shape, generator=g) * (10.0 ** (step - 3))
            p_ref.grad = grad.clone()
            ref.step()
            p_gpu.grad = grad.to(DEV)
            opt.step()
            accum = torch.zeros(ntex, 4, device=DEV)
            accum[:, :C] = grad.reshape(C, ntex).t().to(DEV)
            planar = torch.empty(shape, device=DEV)
            opt_acc.step_from_accum(p_acc, accum, grad_out=planar)
            assert torch.equal(planar.cpu(), grad)
            assert_close(p_gpu, p_ref, f"param after step {step + 1}", rtol=1e-5, atol=1e-6)
            assert_close(p_acc, p_ref, f"param (from accum) after step {step + 1}", rtol=1e-5, atol=1e-6)
        st, rst = opt.state[id(p_gpu)], ref.state[p_ref]
        assert st["step"] == int(rst["step"])
        assert_close(st["exp_avg"], rst["exp_avg"], "exp_avg", rtol=1e-5, atol=1e-7)
        assert_close(st["exp_avg_sq"], rst["exp_avg_sq"], "exp_avg_sq", rtol=1e-5, atol=1e-9)
    with pytest.raises(RuntimeError, match="no CPU path"):
        lp.optim.FusedAdam([torch.zeros(4)])
    a = _lib.LpAdamArgs()
    assert _lib.lib().lp_adam_step(ctypes.byref(a), None) == _lib.LP_ERR_BAD_ARG
