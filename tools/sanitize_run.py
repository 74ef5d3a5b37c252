"""Small end-to-end runs for compute-sanitizer (memcheck / racecheck / synccheck): config 1 (blub, 64 x 64, nearest, latent_paint
Renderer, forward + backward), config 3 geometry (teddy, mesh flavour, 4 views at 64 x 64, normals + lighting), the benched
split chain on a small config-2 scene (DeviceStep: prepare -> raster -> shade -> backward, interleaved gradient + fused
Adam), the fused composition with the bicubic resize, and the bicubic fetch.  Prints 'sanitize_run ok'."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import latent_nerf_test_b200 as lp
from tests.common import latent_paint_views, mesh_views, rnd, scene

DEV = "cuda:0"
verts, faces, uv = scene("blub", 0.6, 0.25)
tex = rnd((1, 4, 128, 128), 1, 0.4).to(DEV).requires_grad_(True)
r = lp.LatentPaintRenderer(DEV, dim=(64, 64), interpolation_mode="nearest")
r.keep_buffers = True
img, mask = r.render_single_view_texture(verts.to(DEV), faces.to(DEV), uv.to(DEV), tex, elev=1.0, azim=0.7, radius=1.25, look_at_height=0.25)
img.backward(rnd(tuple(img.shape), 2).to(DEV))
r.depth_map()
rb = lp.LatentPaintRenderer(DEV, dim=(96, 96), interpolation_mode="bicubic")
img, _ = rb.render_single_view_texture(verts.to(DEV), faces.to(DEV), uv.to(DEV), tex, elev=1.0, azim=0.7, radius=1.25, look_at_height=0.25)
img.backward(rnd(tuple(img.shape), 3).to(DEV))
env = lp.meshio.find_shape("env_sphere")
class M:  # noqa: E701
    def __init__(s, v, f): s.vertices, s.faces = v, f
colors = rnd((1, env.faces.shape[0], 3, 4), 3).to(DEV).requires_grad_(True)
out = lp.textured_mesh.render_train(lp.LatentPaintRenderer(DEV, dim=(96, 96), interpolation_mode="bilinear"), M(verts.to(DEV), faces.to(DEV)),
                                    uv.to(DEV), tex, M(env.vertices.to(DEV), env.faces.to(DEV)), colors, theta=1.0, phi=0.7, radius=1.25)
out["image"].sum().backward()

vt, ft, ut = scene("teddy", 1.0, 0.0)
radius, theta, phi = mesh_views(4, seed=0)
rm = lp.LatentPaintMeshRenderer(DEV, dim=(64, 64))
t2 = rnd((1, 4, 128, 128), 1, 0.4).to(DEV).requires_grad_(True)
outs = rm.render_single_view_texture(vt.to(DEV), ft.to(DEV), ut.to(DEV), t2, theta, phi, radius, dims=(64, 64), is_body=True)
outs[0].backward(rnd(tuple(outs[0].shape), 4).to(DEV))

from bench import DeviceStep, WORKLOADS, workload_cameras
w = dict(WORKLOADS["c2"], B=2, H=160, W=208, T=256)
vn, fn_, un = scene(w["shape"], w["scale"], w["dy"])
geom = (vn.to(DEV).float().contiguous(), fn_.to(DEV, torch.int32).contiguous(), un.to(DEV).float().reshape(-1, 3, 2).contiguous())
st = DeviceStep(geom, w, workload_cameras(w, 2, 5), 1, torch.device(DEV), grad_layout="interleaved")
st.run_split()
p = st.tex.clone()
lp.optim.FusedAdam([p], lr=0.01, betas=(0.9, 0.99), eps=1e-15).step_from_accum(p, st.accum.view(torch.float32).view(-1, 4))
st2 = DeviceStep(geom, w, workload_cameras(w, 2, 5), 1, torch.device(DEV))
st2.fwd.flags |= lp._lib.LP_FLAG_MICRO_ON
st2.run()
torch.cuda.synchronize()
assert float(st.mask.sum()) > 0 and float(outs[1].sum()) > 0
line = __import__("ctypes").c_int32(0)
nfail = lp._lib.lib().lp_check_failures(__import__("ctypes").byref(line))
print(f"device-side invariant checks: {nfail} violations" + (f" (first at lp_b200.cu:{line.value})" if nfail else "") +
      (" [checked build]" if os.environ.get("LP_B200_LIB") else " [ordinary build: checks compiled out]"))
assert nfail == 0
print("sanitize_run ok")
