"""Device timings of the rows added per SURVEY.md §8(f): fused model-level composition (rank 1) and the fused
Adam step (rank 4), each beside the unfused torch sequence it replaces.  Prints one JSON object."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import latent_nerf_test_b200 as lp
from tests.common import rnd, scene

DEV = "cuda:0"
out = {}


def timed(fn, iters=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / iters          # us


class M:
    def __init__(self, v, f): self.vertices, self.faces = v, f

# ---- rank 1: render_train at the reference's training geometry (blub, 64 x 64 latent grid, 4-ch 128^2 texture)
verts, faces, uv = scene("blub", 0.6, 0.25)
env = lp.meshio.find_shape("env_sphere")
obj, envm, uvd = M(verts.to(DEV), faces.to(DEV)), M(env.vertices.to(DEV), env.faces.to(DEV)), uv.to(DEV)
tex = rnd((1, 4, 128, 128), 1, 0.4).to(DEV).requires_grad_(True)
col = torch.rand(1, env.faces.shape[0], 3, 4, device=DEV).requires_grad_(True)
g = rnd((1, 4, 64, 64), 2).to(DEV)
r = lp.LatentPaintRenderer(DEV, dim=(64, 64), interpolation_mode="bilinear")
view = dict(elev=1.0, azim=0.7, radius=1.25, look_at_height=0.25)
def fused():
    tex.grad = None; col.grad = None
    o = lp.textured_mesh.render_train(r, obj, uvd, tex, envm, col, 1.0, 0.7, 1.25, dy=0.25)
    o["image"].backward(g)
def unfused():
    tex.grad = None; col.grad = None
    fg, mask = r.render_single_view_texture(obj.vertices, obj.faces, uvd, tex, **view)
    bg, _ = r.render_single_view(envm, col, **view)
    mask = mask.detach()
    (bg * (1 - mask) + fg * mask).backward(g)
out["render_train_64_us"] = {"fused_composition": round(timed(fused), 1), "two_renders_plus_torch_composition": round(timed(unfused), 1)}

# ---- rank 4: Adam on the config-2 texture (3 x 1024^2), from the interleaved accumulation buffer and from a planar gradient
T, C = 1024, 3
p = torch.nn.Parameter(0.4 * torch.randn(1, C, T, T, device=DEV))
accum = torch.randn(T * T, 4, device=DEV)
opt = lp.optim.FusedAdam([p], lr=0.01, betas=(0.9, 0.99), eps=1e-15)
t_acc = timed(lambda: opt.step_from_accum(p, accum), 100)
p.grad = torch.randn_like(p)
t_planar = timed(opt.step, 100)
q = torch.nn.Parameter(p.detach().clone()); q.grad = torch.randn_like(q)
ref = torch.optim.Adam([q], lr=0.01, betas=(0.9, 0.99), eps=1e-15)
t_torch = timed(ref.step, 100)
ref_fused = torch.optim.Adam([q], lr=0.01, betas=(0.9, 0.99), eps=1e-15, fused=True)
t_torch_fused = timed(ref_fused.step, 100)
bytes_acc = 16 * T * T + 6 * 4 * C * T * T          # accum in; param, exp_avg, exp_avg_sq in and out
bytes_planar = 7 * 4 * C * T * T
peak = 6461.5
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
out["adam_3x1024x1024_us"] = {"k_adam_from_accum(unpack fused)": round(t_acc, 1), "k_adam_planar": round(t_planar, 1),
                              "torch.optim.Adam(foreach)": round(t_torch, 1), "torch.optim.Adam(fused=True)": round(t_torch_fused, 1)}
out["adam_roofline"] = {"bytes_from_accum": bytes_acc, "GBps_from_accum": round(bytes_acc / t_acc / 1e3, 1),
                        "frac_from_accum": round(bytes_acc / t_acc / 1e3 / peak, 3), "bytes_planar": bytes_planar,
                        "frac_planar": round(bytes_planar / t_planar / 1e3 / peak, 3), "peak_GBps": peak}
print(json.dumps(out))
