"""Parity + device timing of the other BASELINE.json configs (c1, c3, c4) on one GPU.
Prints one JSON object; used for DESIGN.md, not a bench line."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import latent_nerf_test_b200 as lp
from oracle import kaolin_shim as kal, renderer_ref
from tests.common import latent_paint_views, mesh_views, rnd, scene

DEV = "cuda:0"
out = {}


def timed(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


# ---- c1: blub, 4ch 128^2 latent texture, 64x64, 1 view, nearest (latent_paint flavour)
verts, faces, uv = scene("blub", 0.6, 0.25)
vd, fd, ud = verts.to(DEV), faces.to(DEV), uv.to(DEV)
tex = rnd((1, 4, 128, 128), 1, 0.4).to(DEV).requires_grad_(True)
r = lp.LatentPaintRenderer(DEV, dim=(64, 64), interpolation_mode="nearest")
g = rnd((1, 4, 64, 64), 2).to(DEV)
def c1():
    tex.grad = None
    img, _ = r.render_single_view_texture(vd, fd, ud, tex, elev=1.0, azim=0.7, radius=1.25, look_at_height=0.25)
    img.backward(g)
out["c1_ms_per_view_fwd_bwd_python_api"] = timed(c1, 50)

# ---- c3: teddy, 4ch 512^2 latent texture, batch 64 (mesh flavour, bilinear), 64x64 and 512x512
verts, faces, uv = scene("teddy", 1.0, 0.0)
vd, fd, ud = verts.to(DEV), faces.to(DEV), uv.to(DEV)
radius, theta, phi = mesh_views(64, seed=0)
rm = lp.LatentPaintMeshRenderer(DEV, dim=(64, 64))
rm.keep_buffers = True
for dims in [(64, 64), (512, 512)]:
    B = 64 if dims[0] == 64 else 8
    tex = rnd((1, 4, 512, 512), 1, 0.4).to(DEV).requires_grad_(True)
    g = rnd((B, 4, dims[1], dims[0]), 2).to(DEV)
    th, ph, ra = theta[:B].to(DEV), phi[:B].to(DEV), radius[:B].to(DEV)
    def c3():
        tex.grad = None
        o = rm.render_single_view_texture(vd, fd, ud, tex, th, ph, ra, dims=dims, is_body=True)
        o[0].backward(g)
    out[f"c3_{dims[0]}_ms_per_step_B{B}"] = timed(c3, 20)
# parity of one c3 batch (8 views, 64x64) against the oracle, cameras from the device kernel fed to the oracle
tex = rnd((1, 4, 512, 512), 1, 0.4).to(DEV).requires_grad_(True)
o = rm.render_single_view_texture(vd, fd, ud, tex, theta[:8], phi[:8], radius[:8], dims=(64, 64), is_body=True)
gg = rnd((8, 4, 64, 64), 3)
o[0].backward(gg.to(DEV))
tc = tex.detach().cpu().clone().requires_grad_(True)
ref = renderer_ref.LatentPaintMeshRendererRef(dim=(64, 64))
ro = ref.render_single_view_texture(verts, faces, uv, tc, theta[:8], phi[:8], radius[:8], dims=(64, 64), is_body=True)
ro[0].backward(gg)
out["c3_face_idx_exact"] = bool(torch.equal(rm.last_buffers["face_idx"].cpu().long(), ref.last["face_idx"]))
out["c3_image_max_abs_err"] = float((o[0].detach().cpu() - ro[0].detach()).abs().max())
out["c3_grad_max_abs_err"] = float((tex.grad.cpu() - tc.grad).abs().max())

# ---- c4: sphere subdivided 5x (1 310 720 faces), 3ch 1024^2 texture, 1024x1024, batch 16
t0 = time.time()
verts, faces, uv = scene("sphere", 0.6, 0.25, subdivide=5)
out["c4_faces"] = int(faces.shape[0])
vd, fd, ud = verts.to(DEV), faces.to(DEV), uv.to(DEV)
tex = rnd((1, 3, 1024, 1024), 1, 0.4).to(DEV).requires_grad_(True)
r4 = lp.LatentPaintRenderer(DEV, dim=(1024, 1024), interpolation_mode="bilinear")
r4.keep_buffers = True
img, mask = r4.render_single_view_texture(vd, fd, ud, tex, elev=1.1, azim=0.3, radius=1.2, look_at_height=0.25)
g1 = rnd((1, 3, 1024, 1024), 2)
img.backward(g1.to(DEV))
tc = tex.detach().cpu().clone().requires_grad_(True)
kal.RASTER_IMPL = "bbox"
ref = renderer_ref.LatentPaintRendererRef(dim=(1024, 1024), interpolation_mode="bilinear")
ri, rmk = ref.render_single_view_texture(verts, faces, uv, tc, elev=1.1, azim=0.3, radius=1.2, look_at_height=0.25)
ri.backward(g1)
out["c4_face_idx_exact"] = bool(torch.equal(r4.last_buffers["face_idx"].cpu().long(), ref.last["face_idx"]))
out["c4_mask_exact"] = bool(torch.equal(mask.cpu(), rmk))
out["c4_image_max_abs_err"] = float((img.detach().cpu() - ri.detach()).abs().max())
out["c4_grad_max_abs_err"] = float((tex.grad.cpu() - tc.grad).abs().max())
out["c4_coverage"] = float(mask.mean())
# device timing of a 16-view batch through the C ABI (bench.DeviceStep)
from bench import DeviceStep, cameras_for, make_views
import ctypes
w = dict(B=16, H=1024, W=1024, C=3, T=1024, interp="bilinear")
geom = (vd.float().contiguous(), fd.to(torch.int32).contiguous(), ud.float().reshape(-1, 3, 2).contiguous())
radius, theta, phi = make_views(16, 0)
st = DeviceStep(geom, w, cameras_for(radius, theta, phi, 0.25), 1, torch.device(DEV))
out["c4_workspace_GB"] = st.ws.numel() / 1e9
out["c4_ms_per_step_B16"] = timed(st.run, 5)
V, F = verts.shape[0], faces.shape[0]
bytes_step = 16 * (12 * V + 36 * F + 1024 * 1024 * (8 * 3 + 20)) + 8 * 3 * 1024 * 1024
out["c4_views_per_s"] = 16 / (out["c4_ms_per_step_B16"] * 1e-3)
out["c4_roofline_frac"] = bytes_step / (out["c4_ms_per_step_B16"] * 1e-3) / 1e9 / 6461.5
print(json.dumps(out))
