"""N-GPU parity of the sharded render path (run under torchrun): every rank renders its shard of the views with the
mesh-flavour Renderer, the texture gradient is all-reduced (NCCL and, when available, the library's in-switch kernel),
and the result is compared with the single-GPU gradient over the whole batch and with the CPU oracle on rank 0."""
import json, os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import latent_nerf_test_b200 as lp
from latent_nerf_test_b200.parallel import GradientBucket, SymmetricGradientBuffer, shard_views
from tests.common import mesh_views, rnd, scene

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
B, dims, T = 12, (96, 96), 256
verts, faces, uv = scene("teddy", 1.0, 0.0)
radius, theta, phi = mesh_views(B, seed=3)
tex0 = rnd((1, 4, T, T), 1, 0.4)
g = rnd((B, 4, dims[1], dims[0]), 2)
r = lp.LatentPaintMeshRenderer(dev, dim=dims)
vd, fd, ud = verts.to(dev), faces.to(dev), uv.to(dev)

def grad_of(lo, hi, tex):
    out = r.render_single_view_texture(vd, fd, ud, tex, theta[lo:hi], phi[lo:hi], radius[lo:hi], dims=dims, is_body=True)
    out[0].backward(g[lo:hi].to(dev))

res = {}
# NCCL through the flat bucket
tex = torch.nn.Parameter(tex0.to(dev))
bucket = GradientBucket([tex]); bucket.zero_()
lo, hi = shard_views(B, rank, world)
if hi > lo: grad_of(lo, hi, tex)
bucket.all_reduce()
sharded = tex.grad.clone()
# the whole batch on this GPU alone
tex1 = tex0.to(dev).requires_grad_(True)
grad_of(0, B, tex1)
full = tex1.grad
res["nccl_max_abs_err_vs_single_gpu"] = float((sharded - full).abs().max())
res["nccl_allclose_1e-4_1e-5"] = bool(torch.allclose(sharded, full, rtol=1e-4, atol=1e-5))
# the library's own exchange over symmetric memory: the backward scatters straight into the symmetric buffer
try:
    sb = SymmetricGradientBuffer(tex0.numel(), dev)
    tex2 = torch.nn.Parameter(tex0.to(dev))
    tex2.grad = sb.view(tuple(tex0.shape)); sb.flat.zero_()
    if hi > lo: grad_of(lo, hi, tex2)
    assert tex2.grad.data_ptr() == sb.flat.data_ptr()
    sb.all_reduce(); torch.cuda.synchronize()
    res["symm_mode"] = sb.mode
    res["symm_max_abs_err_vs_single_gpu"] = float((tex2.grad - full).abs().max())
    res["symm_allclose_1e-4_1e-5"] = bool(torch.allclose(tex2.grad, full, rtol=1e-4, atol=1e-5))
except Exception as exc:
    res["symm_error"] = str(exc)[:200]
if rank == 0:
    from oracle import renderer_ref
    tc = tex0.clone().requires_grad_(True)
    ro = renderer_ref.LatentPaintMeshRendererRef(dim=dims).render_single_view_texture(verts, faces, uv, tc, theta, phi, radius, dims=dims, is_body=True)
    ro[0].backward(g)
    res["nccl_max_abs_err_vs_cpu_oracle"] = float((sharded.cpu() - tc.grad).abs().max())
    res["oracle_grad_abs_max"] = float(tc.grad.abs().max())
    print(json.dumps({"world": world, "views": B, **res}))
dist.barrier(); dist.destroy_process_group()
