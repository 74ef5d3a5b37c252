"""GPU parity of the BENCHED configuration (run with -m gpu): the exact call chain ``bench.py`` times —
``DeviceStep`` through the C ABI with the split pipeline (``lp_render_prepare`` -> ``lp_render_raster`` ->
``lp_render_shade`` -> ``lp_render_backward``), tile flags, vector REDs and ``LP_FLAG_GRAD_OVERWRITE`` /
``LP_FLAG_GRAD_INTERLEAVED`` — against the CPU oracle on the same seeded inputs, at BASELINE.json's full sizes
(configs[1], [2], [3]), plus the depth / barycentric buffers and UVs outside the unit square.

Bars (BASELINE.json north_star): face_idx and the 0/1 mask bit-exact; pixels and gradients rtol 1e-4 / atol 1e-5;
the one texel that sums tens of thousands of background pixels gets the derived accumulation bound of tests/common.py.
"""
import ctypes

import numpy as np
import pytest
import torch

import latent_nerf_test_b200 as lp
from latent_nerf_test_b200 import _lib
from oracle import kaolin_shim as kal
from oracle import renderer_ref
from tests.common import accumulation_terms, assert_close, assert_texture_grad_close, fp64_corner_texel, rnd, scene

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    kal.RASTER_IMPL = "bbox"


def _geom(verts, faces, uv):
    return (verts.to(DEV).float().contiguous(), faces.to(DEV, torch.int32).contiguous(),
            uv.to(DEV).float().reshape(-1, 3, 2).contiguous())


def _oracle_lp_batch(w, verts, faces, uv, tex, grad, views):
    """latent_paint flavour, one view per call as the reference does; → images, masks, face_idx, summed gradient."""
    radius, theta, phi = views
    t = tex.detach().cpu().clone().requires_grad_(True)
    ref = renderer_ref.LatentPaintRendererRef(dim=(w["W"], w["H"]), interpolation_mode=w["interp"])
    images, masks, fidx, depth, bary = [], [], [], [], []
    for i in range(len(theta)):
        image, mask = ref.render_single_view_texture(verts, faces, uv, t, elev=float(theta[i]), azim=float(phi[i]),
                                                     radius=float(radius[i]), look_at_height=w["dy"])
        image.backward(grad[i:i + 1])
        images.append(image.detach()); masks.append(mask); fidx.append(ref.last["face_idx"])
        depth.append(kal.LAST["depth"].clone()); bary.append(kal.LAST["bary"].clone())
    return torch.cat(images), torch.cat(masks), torch.cat(fidx), t.grad[0], torch.cat(depth), torch.cat(bary)


@pytest.mark.parametrize("exchange_form", [False, True, "interleaved"])
def test_config2_benched_step_vs_oracle(exchange_form):
    """configs[1] exactly as bench.py runs it: B = 8, 512 x 512, 3 x 1024^2 texture, split pipeline.  With
    ``exchange_form`` the backward leaves the gradient interleaved and lp_allreduce_unpack (world 1) finishes it —
    the form the N > 1 bench uses."""
    from bench import WORKLOADS, DeviceStep, make_views, workload_cameras
    w = WORKLOADS["c2"]
    verts, faces, uv = scene(w["shape"], w["scale"], w["dy"])
    B, C, T = w["B"], w["C"], w["T"]
    ntex = T * T
    both = torch.zeros(C * ntex + 4 * ntex, device=DEV)
    kw = dict(grad_tex=both[:C * ntex].view(C, T, T), accum=both[C * ntex:]) if exchange_form is True else {}
    if exchange_form == "interleaved":          # bench.py's default at N = 1: the gradient stays (T,T,4) for the fused optimiser
        kw = dict(grad_layout="interleaved")
    st = DeviceStep(_geom(verts, faces, uv), w, workload_cameras(w, B, 0), 1, torch.device(DEV), **kw)
    st.run_split()
    if exchange_form is True:
        ptrs = torch.tensor([both.data_ptr()], dtype=torch.int64, device=DEV)
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(_lib.lib().lp_allreduce_unpack(None, ctypes.c_void_p(ptrs.data_ptr()), 4 * C * ntex, 0, ntex, C, 0, 1, stream))
    torch.cuda.synchronize()
    oi, om, ofi, og, _, _ = _oracle_lp_batch(w, verts, faces, uv, st.tex, st.grad_image.cpu(), make_views(B, 0))
    assert torch.equal(st.mask.cpu(), om), "0/1 mask (= face_idx > -1) must be bit-exact"
    assert torch.equal(st.mask.cpu()[:, 0] > 0, ofi >= 0)
    assert_close(st.image, oi, "image of the benched step")
    assert_close(st.gradient(), og, "texture gradient of the benched step (8 views summed)")
    # coverage flags: 1 exactly where the 8 x 4-pixel footprint holds a covered pixel
    cov = (ofi >= 0).reshape(B, w["H"] // 4, 4, w["W"] // 8, 8).any(dim=4).any(dim=2)
    assert torch.equal(st.footprint_any.cpu().bool(), cov)
    # second run of the same buffers: deterministic visibility, gradient overwritten (not accumulated twice)
    img1, g1 = st.image.clone(), st.gradient().clone()
    st.run_split()
    if exchange_form is True:
        _lib.check(_lib.lib().lp_allreduce_unpack(None, ctypes.c_void_p(ptrs.data_ptr()), 4 * C * ntex, 0, ntex, C, 0, 1, stream))
    torch.cuda.synchronize()
    assert torch.equal(st.image, img1)
    # (the scatter's atomic order differs from run to run: same terms, another summation order)
    assert_close(st.gradient(), g1, "gradient of a second step on the same buffers")
    if exchange_form == "interleaved":
        # ... and the fused optimiser consumes exactly that buffer: one Adam step from it equals torch's on the oracle gradient
        p_gpu = st.tex.clone()
        opt = lp.optim.FusedAdam([p_gpu], lr=0.01, betas=(0.9, 0.99), eps=1e-15)
        opt.step_from_accum(p_gpu, st.accum.view(torch.float32).view(-1, 4))
        p_ref = torch.nn.Parameter(st.tex.detach().cpu().clone())
        p_ref.grad = og[None].clone()
        torch.optim.Adam([p_ref], lr=0.01, betas=(0.9, 0.99), eps=1e-15).step()
        moved = (p_ref.detach() - st.tex.cpu()).abs() > 1e-3           # texels that received gradient
        assert int(moved.sum()) > 100000
        # Adam's first step is lr * sign(g) wherever |g| >> eps: compare where the gradient is not within rounding of zero
        solid = og[None].abs() > 1e-4
        assert_close(p_gpu.cpu()[solid], p_ref.detach()[solid], "texture after one fused Adam step from the interleaved gradient", rtol=1e-5, atol=1e-6)


def test_accumulation_buffer_cleared_next_to_the_texture_fetch():
    """bench.py's pipelined step clears the backward's accumulation buffer on another stream, next to the texture fetch,
    and passes LP_FLAG_GRAD_NO_CLEAR: same gradient as the call that clears the buffer itself, from eager launches and
    from a CUDA graph (the fork / join of the clear is captured); and the flag alone really skips the clear (a second
    backward accumulates on top of the first)."""
    from bench import WORKLOADS, DeviceStep, workload_cameras
    w = dict(WORKLOADS["c2"], B=3, H=160, W=208, T=256)
    verts, faces, uv = scene(w["shape"], w["scale"], w["dy"])
    dev = torch.device(DEV)
    geom = _geom(verts, faces, uv)
    ref = DeviceStep(geom, w, workload_cameras(w, w["B"], 4), 1, dev, grad_layout="interleaved")
    st = DeviceStep(geom, w, workload_cameras(w, w["B"], 4), 1, dev, grad_layout="interleaved")
    main, aux = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    h = ctypes.c_void_p(main.cuda_stream)
    ref.run_split()
    torch.cuda.synchronize()
    g_ref = ref.gradient().clone()
    assert float(g_ref.abs().sum()) > 0
    with torch.cuda.stream(main):
        st.accum.fill_(77)                               # whatever a previous step left behind
        st.prepare(h, True)
        st.shade_backward(h, main, True, aux=aux)
    torch.cuda.synchronize()
    assert_close(st.gradient(), g_ref, "gradient with the buffer cleared on the side stream")
    assert torch.equal(st.image, ref.image)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(main):
        with torch.cuda.graph(g, stream=main):
            st.prepare(h, True)
            st.shade_backward(h, main, True, aux=aux)
        st.accum.fill_(5)
        g.replay()
        g.replay()
    torch.cuda.synchronize()
    assert_close(st.gradient(), g_ref, "gradient from the captured fork / clear / join")
    # the flag alone: no clear, so a second backward adds to the first
    st.bwd.flags |= _lib.LP_FLAG_GRAD_NO_CLEAR
    with torch.cuda.stream(main):
        _lib.check(_lib.lib().lp_render_backward(ctypes.byref(st.bwd), h))
    torch.cuda.synchronize()
    assert_close(st.gradient(), 2 * g_ref, "two backwards without a clear in between", rtol=1e-4, atol=2e-5)


def test_config2_visibility_buffers_vs_oracle():
    """face_idx, depth and barycentric weights of configs[1] (B = 8 in one call) against the oracle's buffers:
    same fp32 expression tree in the same order, so all three are compared bit for bit."""
    from bench import WORKLOADS, DeviceStep, make_views, workload_cameras
    w = WORKLOADS["c2"]
    verts, faces, uv = scene(w["shape"], w["scale"], w["dy"])
    B, H, W = w["B"], w["H"], w["W"]
    st = DeviceStep(_geom(verts, faces, uv), w, workload_cameras(w, B, 0), 1, torch.device(DEV))
    face_idx = torch.empty(B, H, W, dtype=torch.int32, device=DEV)
    bary = torch.empty(B, H, W, 3, device=DEV)
    depth = torch.empty(B, H, W, device=DEV)
    st.fwd.face_idx, st.fwd.bary, st.fwd.depth = face_idx.data_ptr(), bary.data_ptr(), depth.data_ptr()
    st.run_split()
    torch.cuda.synchronize()
    oi, om, ofi, og, odepth, obary = _oracle_lp_batch(w, verts, faces, uv, st.tex, st.grad_image.cpu(), make_views(B, 0))
    assert torch.equal(face_idx.cpu().long(), ofi)
    assert torch.equal(depth.cpu(), odepth), f"depth: max |diff| {float((depth.cpu() - odepth).abs().max()):.3e}"
    assert torch.equal(bary.cpu(), obary), f"bary: max |diff| {float((bary.cpu() - obary).abs().max()):.3e}"
    assert_close(st.image, oi, "image (buffers requested)")
    assert_close(st.grad_tex, og, "texture gradient (buffers requested)")


def test_config3_mesh_flavour_B64_vs_oracle():
    """configs[2]: teddy, 4 x 512^2 latent texture, 64 views at 64 x 64, latent_paint_mesh flavour, through the
    benched DeviceStep chain.  The image is NOT masked in this flavour, so every uncovered pixel of every view
    back-propagates into texel (T-1, 0): that texel is compared with the fp64 sum under the derived bound."""
    from bench import WORKLOADS, DeviceStep, make_views, workload_cameras, LOOK_AT_BODY
    w = WORKLOADS["c3"]
    verts, faces, uv = scene(w["shape"], w["scale"], w["dy"])
    B, H, W, C, T = w["B"], w["H"], w["W"], w["C"], w["T"]
    st = DeviceStep(_geom(verts, faces, uv), w, workload_cameras(w, B, 0), 1, torch.device(DEV))
    face_idx = torch.empty(B, H, W, dtype=torch.int32, device=DEV)
    st.fwd.face_idx = face_idx.data_ptr()
    st.run_split()
    torch.cuda.synchronize()
    radius, theta, phi = make_views(B, 0, "mesh")
    t = st.tex.detach().cpu().clone().requires_grad_(True)
    ref = renderer_ref.LatentPaintMeshRendererRef(dim=(W, H))
    routs = ref.render_single_view_texture(verts, faces, uv, t, theta, phi, radius, dims=(W, H), is_body=True)
    g = st.grad_image.cpu()
    routs[0].backward(g)
    ofi = ref.last["face_idx"]
    assert torch.equal(face_idx.cpu().long(), ofi), "face_idx must be bit-exact"
    assert torch.equal(st.mask.cpu()[:, 0] > 0, ofi >= 0)
    for got, want, name in ((st.image, routs[0], "image"), (st.mask, routs[1], "mask"), (st.normals, routs[2], "normals"),
                            (st.lighting, routs[3], "lighting")):
        assert_close(got, want, name)
    n, rms, ref64 = fp64_corner_texel(ref.last["uv"], t, g, ofi)
    assert n > 10000, "the background texel must be a real accumulation in this test"
    assert_texture_grad_close(st.grad_tex, t.grad[0], "texture gradient (64 views)", background=(n, rms, ref64),
                              terms=accumulation_terms(ref.last["uv"], g, t.shape))


def test_config4_full_size_one_view_vs_oracle():
    """configs[3] geometry at full size: sphere.obj subdivided five times (1 310 720 faces), 1024 x 1024,
    3 x 1024^2 texture, one view, through the C ABI chain of the bench (micro-face path)."""
    from bench import WORKLOADS, DeviceStep, make_views, workload_cameras
    w = dict(WORKLOADS["c4"], B=1)
    verts, faces, uv = scene(w["shape"], w["scale"], w["dy"], subdivide=w["subdivide"])
    assert faces.shape[0] == 1310720
    H, W = w["H"], w["W"]
    st = DeviceStep(_geom(verts, faces, uv), w, workload_cameras(w, 1, 3), 1, torch.device(DEV))
    face_idx = torch.empty(1, H, W, dtype=torch.int32, device=DEV)
    st.fwd.face_idx = face_idx.data_ptr()
    st.run_split()
    torch.cuda.synchronize()
    oi, om, ofi, og, _, _ = _oracle_lp_batch(w, verts, faces, uv, st.tex, st.grad_image.cpu(), make_views(1, 3))
    assert torch.equal(face_idx.cpu().long(), ofi)
    assert torch.equal(st.mask.cpu(), om)
    assert float(om.mean()) > 0.2
    assert_close(st.image, oi, "image")
    assert_close(st.grad_tex, og, "texture gradient")


@pytest.mark.parametrize("mode", ["nearest", "bilinear"])
def test_uvs_outside_the_unit_square(mode):
    """Meshes whose vt lie outside [0,1] (animal / hand / potion.obj in the reference's shapes/): kaolin clamps inside
    texture_mapping, the pixel stays covered (mask 1) and its gradient goes to the border texel.  Fused and split
    forward, autograd and C-ABI backward must all agree with the oracle."""
    from bench import DeviceStep, make_views, cameras_for
    verts, faces, uv = scene("blub", 0.6, 0.25)
    uv = uv * 2.0 - 0.5                                          # u, v in [-0.5, 1.5]: a third of the surface is outside
    C, T, H, W = 4, 64, 96, 80
    tex = rnd((1, C, T, T), 1, 0.4)
    g = rnd((1, C, H, W), 2)
    view = dict(elev=1.0, azim=0.7, radius=1.25, look_at_height=0.25)
    tc = tex.clone().requires_grad_(True)
    ref = renderer_ref.LatentPaintRendererRef(dim=(W, H), interpolation_mode=mode)
    oi, om = ref.render_single_view_texture(verts, faces, uv, tc, **view)
    oi.backward(g)
    ouv = ref.last["uv"]
    assert float((ouv[..., 0] < 0).float().mean()) > 0.01, "the test needs covered pixels with negative u"
    # fused forward + autograd backward (Python API)
    tg = tex.to(DEV).requires_grad_(True)
    r = lp.LatentPaintRenderer(DEV, dim=(W, H), interpolation_mode=mode)
    r.keep_buffers = True
    ig, mg = r.render_single_view_texture(verts.to(DEV), faces.to(DEV), uv.to(DEV), tg, **view)
    ig.backward(g.to(DEV))
    assert torch.equal(mg.cpu(), om)
    assert_close(ig, oi, f"image, fused ({mode})")
    assert_close(tg.grad, tc.grad, f"grad_texture, fused ({mode})")
    assert_close(r.last_buffers["uv"], ouv, "uv returned to the caller (unclamped, 0 where uncovered)")
    # split forward + C-ABI backward (the benched chain)
    w = dict(B=1, H=H, W=W, C=C, T=T, interp=mode, flavour="lp", dy=0.25)
    cams = lp.camera.camera_from_view(torch.tensor(view["elev"]), torch.tensor(view["azim"]), view["radius"], 0.25)
    st = DeviceStep(_geom(verts, faces, uv), w, cams, 1, torch.device(DEV))
    st.tex.copy_(tex)
    st.grad_image.copy_(g)
    st.run_split()
    torch.cuda.synchronize()
    assert torch.equal(st.mask.cpu(), om)
    assert_close(st.image, oi, f"image, split ({mode})")
    assert_close(st.grad_tex, tc.grad[0], f"grad_texture, split ({mode})")


@pytest.mark.parametrize("pdl", [1, 0])
def test_bin_overflow_cascade_and_rerun(pdl):
    """Fixed-capacity cells: with the micro-face path switched off, blub's 14 208 faces at 64 x 64 put thousands of
    faces into single cells, so insertions cascade through the parents up to the root list.  Visibility must not
    care (bit-exact), with and without programmatic dependent launch, and the prepared bins can be rasterized twice."""
    from bench import DeviceStep
    verts, faces, uv = scene("blub", 0.6, 0.25)
    H = W = 64
    w = dict(B=2, H=H, W=W, C=4, T=128, interp="bilinear", flavour="lp", dy=0.25)
    cams = torch.cat([lp.camera.camera_from_view(torch.tensor(e), torch.tensor(a), r, 0.25)
                      for e, a, r in ((1.0, 0.7, 1.25), (2.0, 4.0, 1.05))])
    L = _lib.lib()
    _lib.check(L.lp_set_option(_lib.LP_OPT_PDL, pdl))
    try:
        for micro_flag in (_lib.LP_FLAG_MICRO_OFF, _lib.LP_FLAG_MICRO_ON, 0):
            st = DeviceStep(_geom(verts, faces, uv), w, cams, 1, torch.device(DEV))
            st.fwd.flags |= micro_flag
            face_idx = torch.empty(2, H, W, dtype=torch.int32, device=DEV)
            st.fwd.face_idx = face_idx.data_ptr()
            st.run_split()
            torch.cuda.synchronize()
            ofi = []
            t = st.tex.detach().cpu().clone().requires_grad_(True)
            ref = renderer_ref.LatentPaintRendererRef(dim=(W, H), interpolation_mode="bilinear")
            imgs = []
            for i, (e, a, r) in enumerate(((1.0, 0.7, 1.25), (2.0, 4.0, 1.05))):
                img, _ = ref.render_single_view_texture(verts, faces, uv, t, elev=e, azim=a, radius=r, look_at_height=0.25)
                img.backward(st.grad_image.cpu()[i:i + 1])
                imgs.append(img.detach()); ofi.append(ref.last["face_idx"])
            assert torch.equal(face_idx.cpu().long(), torch.cat(ofi)), f"micro flag {micro_flag:#x}"
            assert_close(st.image, torch.cat(imgs), "image")
            assert_close(st.grad_tex, t.grad[0], "texture gradient")
            # rasterize the same prepared bins again (the tile kernel rewinds its own ticket counter)
            first = face_idx.clone()
            face_idx.fill_(-5)
            stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(L.lp_render_raster(ctypes.byref(st.fwd), stream))
            torch.cuda.synchronize()
            assert torch.equal(face_idx, first)
    finally:
        _lib.check(L.lp_set_option(_lib.LP_OPT_PDL, 1))


def test_decree_switches_flip_kernel_and_oracle_together():
    """The open points of the kaolin restatement are runtime switches in the kernels, in oracle/raster_ref.c and in
    oracle/kaolin_shim.py.  Each one, flipped in both places, must keep face_idx / depth / barycentrics bit-identical
    and the pixels within tolerance — and must actually change something."""
    verts, faces, uv = scene("blub", 0.6, 0.25)
    tex = rnd((1, 4, 64, 64), 1, 0.4).to(DEV)
    view = dict(elev=1.0, azim=0.7, radius=1.25, look_at_height=0.25)
    base = {}
    try:
        for name, attr in (("decree", None), ("half_open", "BBOX_HALF_OPEN"), ("plain_eps", "PLAIN_EPS"), ("affine", "AFFINE_INTERP")):
            r = lp.LatentPaintRenderer(DEV, dim=(160, 128), interpolation_mode="bilinear")
            r.keep_buffers = True
            r.bbox_half_open, r.plain_eps, r.affine_interpolation = name == "half_open", name == "plain_eps", name == "affine"
            kal.BBOX_HALF_OPEN, kal.PLAIN_EPS, kal.AFFINE_INTERP = r.bbox_half_open, r.plain_eps, r.affine_interpolation
            image, mask = r.render_single_view_texture(verts.to(DEV), faces.to(DEV), uv.to(DEV), tex, **view)
            ref = renderer_ref.LatentPaintRendererRef(dim=(160, 128), interpolation_mode="bilinear")
            oi, om = ref.render_single_view_texture(verts, faces, uv, tex.cpu(), **view)
            assert torch.equal(r.last_buffers["face_idx"].cpu().long(), ref.last["face_idx"]), name
            assert torch.equal(r.last_buffers["depth"].cpu(), kal.LAST["depth"]), name
            assert torch.equal(r.last_buffers["bary"].cpu(), kal.LAST["bary"]), name
            assert_close(image, oi, f"image ({name})")
            base[name] = (r.last_buffers["depth"].clone(), r.last_buffers["bary"].clone())
        assert not torch.equal(base["affine"][0], base["decree"][0]) and not torch.equal(base["affine"][1], base["decree"][1])
        # screen-space barycentrics sum to one up to rounding, like the perspective-correct ones
        cov = r.last_buffers["face_idx"] >= 0
        assert_close(base["affine"][1].sum(-1)[cov], torch.ones(int(cov.sum())), "affine barycentrics sum to one", atol=1e-5)

        kal.BBOX_HALF_OPEN = kal.PLAIN_EPS = kal.AFFINE_INTERP = False
        # half-open box: a triangle whose right-most / lowest vertices sit exactly on pixel centres (kaolin-level entry,
        # already projected vertices: W = H = 8 puts pixel centres at multiples of 0.125 + 0.0625... in NDC * 1000 = exact)
        kc = lp.kaolin_compat.make_module()
        fvi = torch.tensor([[[[-0.375, 0.375], [0.375, 0.375], [0.375, -0.375]]]])      # pixel centres (col 2, row 2), (5, 2), (5, 5)
        fvz = torch.full((1, 1, 3), -1.0)
        feat = torch.ones(1, 1, 3, 1)
        got = {}
        for ho in (False, True):
            lp.kaolin_compat.BBOX_HALF_OPEN = kal.BBOX_HALF_OPEN = ho
            _, idx = kc.render.mesh.rasterize(8, 8, fvz.to(DEV), fvi.to(DEV), feat.to(DEV))
            _, oidx = kal.rasterize(8, 8, fvz, fvi, feat)
            assert torch.equal(idx.cpu(), oidx), f"half_open={ho}"
            got[ho] = int((idx >= 0).sum())
        assert got[True] < got[False], "the half-open box must drop the pixels on the box's far edges"
        lp.kaolin_compat.BBOX_HALF_OPEN = kal.BBOX_HALF_OPEN = False

        # SH band-1 order: mesh flavour lighting with a light that tells x from y
        vm, fm, um = scene("sphere", 1.0, 0.0)
        lights = torch.tensor([0.5, 0.9, 0.0, 0.1, 0.0, 0.0, 0.0, 0.0, 0.0])
        th, ph, ra = torch.tensor([1.2, 1.6]), torch.tensor([0.3, 2.0]), torch.tensor([1.8, 2.0])
        out = {}
        for xzy in (False, True):
            rm = lp.LatentPaintMeshRenderer(DEV, dim=(64, 64), lights=lights)
            rm.sh_band1_xzy = xzy
            kal.SH_BAND1_AXES = (0, 2, 1) if xzy else (1, 2, 0)
            o = rm.render_single_view_texture(vm.to(DEV), fm.to(DEV), um.to(DEV), tex, th, ph, ra, dims=(64, 64))
            ro = renderer_ref.LatentPaintMeshRendererRef(dim=(64, 64), lights=lights).render_single_view_texture(
                vm, fm, um, tex.cpu(), th, ph, ra, dims=(64, 64))
            assert_close(o[3], ro[3], f"lighting (band-1 xzy={xzy})")
            out[xzy] = o[3].clone()
        assert float((out[True] - out[False]).abs().max()) > 1e-2
    finally:
        kal.BBOX_HALF_OPEN = kal.PLAIN_EPS = kal.AFFINE_INTERP = False
        kal.SH_BAND1_AXES = (1, 2, 0)
        lp.kaolin_compat.BBOX_HALF_OPEN = False


def test_bicubic_texture_fetch_vs_oracle():
    """The third interpolation mode the reference's Renderer accepts (render.py:9, used at :64): ATen's bicubic
    grid_sample (A = -0.75, unclipped coordinate, per-tap border clamp), forward and backward, through the Renderer
    (blub: UVs inside the unit square; and UVs pushed outside it so taps clamp at the border) and through the operator API."""
    verts, faces, uv = scene("blub", 0.6, 0.25)
    view = dict(elev=1.0, azim=0.7, radius=1.25, look_at_height=0.25)
    for uvs, C, T, dims, white in ((uv, 4, 128, (64, 64), False), (uv * 1.6 - 0.3, 3, 37, (120, 88), True)):
        tex = rnd((1, C, T, T), 1, 0.4)
        g = rnd((1, C, dims[1], dims[0]), 2)
        tg = tex.to(DEV).requires_grad_(True)
        r = lp.LatentPaintRenderer(DEV, dim=dims, interpolation_mode="bicubic")
        r.keep_buffers = True
        ig, mg = r.render_single_view_texture(verts.to(DEV), faces.to(DEV), uvs.to(DEV), tg, white_background=white, **view)
        ig.backward(g.to(DEV))
        tc = tex.clone().requires_grad_(True)
        ref = renderer_ref.LatentPaintRendererRef(dim=dims, interpolation_mode="bicubic")
        oi, om = ref.render_single_view_texture(verts, faces, uvs, tc, white_background=white, **view)
        oi.backward(g)
        assert torch.equal(r.last_buffers["face_idx"].cpu().long(), ref.last["face_idx"]) and torch.equal(mg.cpu(), om)
        # sixteen taps with weights in [-0.07, 0.6] whose magnitudes sum to at most 1.5625^2 (see the resize test)
        assert_close(ig, oi, "bicubic image", rtol=1e-4, atol=1e-5 * 1.5625 ** 2)
        assert_close(tg.grad, tc.grad, "bicubic texture gradient", rtol=1e-4, atol=1e-5 * 1.5625 ** 2)
    kc = lp.kaolin_compat.make_module()
    coords = torch.rand(2, 40, 24, 2, generator=torch.Generator().manual_seed(3)) * 1.2 - 0.1
    maps = rnd((2, 3, 21, 17), 4)
    mg = maps.to(DEV).requires_grad_(True)
    out = kc.render.mesh.texture_mapping(coords.to(DEV), mg, mode="bicubic")
    go = rnd(tuple(out.shape), 5)
    out.backward(go.to(DEV))
    mc = maps.clone().requires_grad_(True)
    oo = kal.texture_mapping(coords, mc, mode="bicubic")
    oo.backward(go)
    assert_close(out, oo, "texture_mapping bicubic", rtol=1e-4, atol=1e-5 * 1.5625 ** 2)
    assert_close(mg.grad, mc.grad, "texture_mapping bicubic gradient", rtol=1e-4, atol=1e-5 * 1.5625 ** 2)


def test_fused_bicubic_resize_and_depth_map():
    """SURVEY.md 8 f ranks 1 and 3.  (1) ``lp_resize_bicubic``: the four bicubic ``F.interpolate(x, (64, 64))`` calls of
    ``TexturedMeshModel.render_train`` (reference src/latent_paint/models/textured_mesh.py:214-218) as one launch, forward
    and backward, against torch on the CPU.  (2) ``renderer.depth_map()``: the (B,1,64,64) min-max normalised depth input
    of depth-conditioned guidance (src/stable_diffusion_depth.py:302-319) against the same steps on the oracle's depth."""
    import torch.nn.functional as F
    xs = [rnd((2, c, 96, 80), 10 + c) for c in (1, 4, 4, 3)]
    gs = [rnd((2, c, 64, 64), 20 + c) for c in (1, 4, 4, 3)]
    xg = [x.to(DEV).requires_grad_(i != 0) for i, x in enumerate(xs)]                  # the mask carries no gradient
    ys = lp.functional.resize_bicubic(xg, (64, 64))
    sum((y * g.to(DEV)).sum() for y, g in zip(ys, gs)).backward()
    xc = [x.clone().requires_grad_(i != 0) for i, x in enumerate(xs)]
    yc = [F.interpolate(x, (64, 64), mode="bicubic") for x in xc]
    sum((y * g).sum() for y, g in zip(yc, gs)).backward()
    for i in range(4):
        assert_close(ys[i], yc[i], f"resized tensor {i}", rtol=1e-4, atol=1e-5 * 1.5625 ** 2)
        if i:
            assert_close(xg[i].grad, xc[i].grad, f"gradient through the resize {i}", rtol=1e-4, atol=1e-5 * 4)
    assert xg[0].grad is None
    up = lp.functional.resize_bicubic([xs[1].to(DEV)], (200, 100))[0]                # enlarging works the same way
    assert_close(up, F.interpolate(xs[1], (200, 100), mode="bicubic"), "upsampled", rtol=1e-4, atol=1e-5 * 1.5625 ** 2)

    verts, faces, uv = scene("teddy", 1.0, 0.0)
    radius, theta, phi = mesh_views_local(3)
    tex = rnd((1, 4, 64, 64), 1).to(DEV)
    for dims in ((64, 64), (128, 96)):
        r = lp.LatentPaintMeshRenderer(DEV, dim=dims)
        r.keep_buffers = True
        r.render_single_view_texture(verts.to(DEV), faces.to(DEV), uv.to(DEV), tex, theta, phi, radius, dims=dims)
        dm = r.depth_map(size=64)
        assert dm.shape == (3, 1, 64, 64) and abs(float(dm.min()) + 1) < 1e-6 and abs(float(dm.max()) - 1) < 1e-6
        ref = renderer_ref.LatentPaintMeshRendererRef(dim=dims)
        ref.render_single_view_texture(verts, faces, uv, tex.cpu(), theta, phi, radius, dims=dims)
        z = kal.LAST["depth"]
        assert torch.equal(r.last_buffers["depth"].cpu(), z)
        inv = torch.where(z < 0, -1.0 / z.clamp(max=-1e-12), torch.zeros_like(z))[:, None]
        if dims != (64, 64):
            inv = F.interpolate(inv, size=(64, 64), mode="bicubic", align_corners=False)   # stable_diffusion_depth.py:310
        want = 2.0 * (inv - inv.min()) / (inv.max() - inv.min()) - 1.0                      # :313
        assert_close(dm, want, f"depth map {dims}", rtol=1e-4, atol=2e-5)


def mesh_views_local(B):
    from tests.common import mesh_views
    return mesh_views(B, seed=11)
