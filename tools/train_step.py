"""BASELINE.json configs[4]: the end-to-end Latent-Paint-mesh SDS training step with the new renderer in the loop.

A slimmed copy of the reference loop (``/root/reference/src/latent_paint_mesh/training/trainer.py:353-407`` zero_grad ->
``train_render_text`` -> step, ``:565-658`` render -> ``diffusion.train_step`` -> ``pred_rgb.backward(gradient=grad)``;
guidance ``src/stable_diffusion.py:248-334``): per step and per rank

    views  ~ the reference's ranges (train_config.py:18-22), ``is_body`` alternating per batch (views_dataset.py:83)
    render   LatentPaintMeshRenderer.render_single_view_texture(verts, faces, uv, texture_img, thetas, phis, rs,
             dims=(64, 64))                      -> pred_rgb (B,4,64,64) latent image          [this repo's kernels]
    guide    noise, add_noise(t), UNet(cat[x_t]*2, t, text_z) (no grad), classifier-free guidance 100,
             grad = w(t) (eps_hat - eps)                                                       [PyTorch, out of scope]
    backward pred_rgb.backward(gradient=grad)    -> texture_img.grad                           [this repo's kernels]
    exchange sum of the texture gradient over the ranks (symmetric-memory kernels, NCCL fallback)
    step     Adam(lr 5e-3, betas (0.9, 0.99), eps 1e-15) on the texture (lp_adam_step)

The UNet is a random-init conditional UNet in plain torch (diffusers is not installed here, and the guidance network is
out of scope anyway): it stands for the cost and the data flow of the SD UNet, not for its weights.  Views shard over
the ranks (``--views`` per rank), weak scaling.  Prints one JSON object on rank 0: steps/s, views/s and the share of
the step spent in the renderer (forward + backward + exchange + optimiser) against the guidance network.

    python tools/train_step.py [--views 8] [--steps 20] [--unet-width 320]
    python -m torch.distributed.run --nproc-per-node 8 ... tools/train_step.py
"""
import argparse
import json
import math
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import latent_nerf_test_b200 as lp  # noqa: E402
from latent_nerf_test_b200.parallel import GradientBucket, SymmetricGradientBuffer  # noqa: E402


# ----------------------------------------------------------------------------- guidance stand-in (PyTorch, out of scope)
class ResBlock(nn.Module):
    def __init__(self, cin, cout, temb):
        super().__init__()
        self.n1, self.c1 = nn.GroupNorm(32, cin), nn.Conv2d(cin, cout, 3, padding=1)
        self.t = nn.Linear(temb, cout)
        self.n2, self.c2 = nn.GroupNorm(32, cout), nn.Conv2d(cout, cout, 3, padding=1)
        self.skip = nn.Conv2d(cin, cout, 1) if cin != cout else nn.Identity()

    def forward(self, x, t):
        h = self.c1(F.silu(self.n1(x))) + self.t(F.silu(t))[:, :, None, None]
        return self.c2(F.silu(self.n2(h))) + self.skip(x)


class CrossAttention(nn.Module):
    def __init__(self, c, ctx=768, heads=8):
        super().__init__()
        self.norm, self.heads = nn.GroupNorm(32, c), heads
        self.q, self.k, self.v, self.o = nn.Linear(c, c, bias=False), nn.Linear(ctx, c, bias=False), nn.Linear(ctx, c, bias=False), nn.Linear(c, c)

    def forward(self, x, ctx):
        B, C, H, W = x.shape
        h = self.norm(x).flatten(2).transpose(1, 2)
        q, k, v = self.q(h), self.k(ctx), self.v(ctx)
        split = lambda t: t.reshape(B, -1, self.heads, C // self.heads).transpose(1, 2)
        a = F.scaled_dot_product_attention(split(q), split(k), split(v)).transpose(1, 2).reshape(B, H * W, C)
        return x + self.o(a).transpose(1, 2).reshape(B, C, H, W)


class TinyConditionalUNet(nn.Module):
    """4 -> 4 channels on 64 x 64 latents, three resolutions, text cross-attention at every level, timestep embedding:
    the shape of SD's UNet2DConditionModel at a fraction of its width."""

    def __init__(self, width=320):
        super().__init__()
        w, temb = width, 4 * width
        self.temb = nn.Sequential(nn.Linear(w, temb), nn.SiLU(), nn.Linear(temb, temb))
        self.inp = nn.Conv2d(4, w, 3, padding=1)
        self.d1, self.a1 = ResBlock(w, w, temb), CrossAttention(w)
        self.d2, self.a2 = ResBlock(w, 2 * w, temb), CrossAttention(2 * w)
        self.mid, self.am = ResBlock(2 * w, 4 * w, temb), CrossAttention(4 * w)
        self.u2, self.b2 = ResBlock(4 * w + 2 * w, 2 * w, temb), CrossAttention(2 * w)
        self.u1, self.b1 = ResBlock(2 * w + w, w, temb), CrossAttention(w)
        self.out = nn.Sequential(nn.GroupNorm(32, w), nn.SiLU(), nn.Conv2d(w, 4, 3, padding=1))
        self.width = w

    def forward(self, x, t, encoder_hidden_states):
        half = self.width // 2
        freqs = torch.exp(-math.log(10000.0) * torch.arange(half, device=x.device) / half)
        e = t.float()[:, None] * freqs[None]
        te = self.temb(torch.cat([e.sin(), e.cos()], dim=1).to(x.dtype)).expand(x.shape[0], -1)
        ctx = encoder_hidden_states
        h1 = self.a1(self.d1(self.inp(x), te), ctx)
        h2 = self.a2(self.d2(F.avg_pool2d(h1, 2), te), ctx)
        m = self.am(self.mid(F.avg_pool2d(h2, 2), te), ctx)
        u2 = self.b2(self.u2(torch.cat([F.interpolate(m, scale_factor=2.0), h2], 1), te), ctx)
        u1 = self.b1(self.u1(torch.cat([F.interpolate(u2, scale_factor=2.0), h1], 1), te), ctx)
        return self.out(u1)


class Guidance:
    """``StableDiffusion.train_step`` of the reference (src/stable_diffusion.py:248-334) in latent mode: returns the SDS
    gradient ``w(t) (eps_hat - eps)`` for ``pred_rgb.backward(gradient=...)``."""

    def __init__(self, device, width, dtype=torch.bfloat16, seed=0):
        torch.manual_seed(seed)
        self.unet = TinyConditionalUNet(width).to(device=device, dtype=dtype).eval()
        betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float32) ** 2      # SD's scaled-linear schedule
        self.alphas = torch.cumprod(1.0 - betas, dim=0).to(device)
        self.min_step, self.max_step = 20, 980
        self.device, self.dtype = device, dtype
        g = torch.Generator(device="cpu").manual_seed(seed + 1)
        self.text_z = torch.randn(2, 77, 768, generator=g).to(device=device, dtype=dtype)         # [uncond, cond]
        self.gen = torch.Generator(device=device).manual_seed(seed + 2)

    @torch.no_grad()
    def train_step(self, latents, guidance_scale=100.0):
        B = latents.shape[0]
        t = torch.randint(self.min_step, self.max_step + 1, [1], dtype=torch.long, device=self.device, generator=self.gen)
        noise = torch.randn(latents.shape, device=self.device, generator=self.gen)
        a = self.alphas[t]
        noisy = a.sqrt() * latents + (1 - a).sqrt() * noise
        ctx = self.text_z.repeat_interleave(B, dim=0)
        pred = self.unet(torch.cat([noisy] * 2).to(self.dtype), t, ctx).float()
        uncond, text = pred.chunk(2)
        pred = uncond + guidance_scale * (text - uncond)
        return a ** 0.5 * (1 - a) * (pred - noise)


# ----------------------------------------------------------------------------- the loop
def views_for(step, rank, B, device):
    """radius, theta, phi in the reference's order and ranges (views_dataset.py:22-24, train_config.py:18-22)."""
    g = torch.Generator().manual_seed(100003 * step + rank)
    radius = torch.rand(B, generator=g) * 1.0 + 1.4
    theta = torch.deg2rad(torch.rand(B, generator=g) * 50.0 + 60.0)
    phi = torch.deg2rad(torch.rand(B, generator=g) * 360.0)
    return theta.to(device), phi.to(device), radius.to(device)


def run(views=8, steps=20, warmup=3, unet_width=320, texture=512, shape="teddy", exchange="auto", quiet=False):
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    dist = None
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=device)
    m = lp.meshio.find_shape(shape)
    verts = lp.meshio.normalize_vertices(m.vertices, 1.0, 0.0).to(device)
    faces, uv = m.faces.to(device), lp.meshio.face_uv_attributes(m).to(device)
    renderer = lp.LatentPaintMeshRenderer(device, dim=(64, 64), interpolation_mode="bilinear")
    C, T = 4, texture
    tex0 = 0.4 * torch.randn(1, C, T, T, generator=torch.Generator().manual_seed(1))
    # the texture's gradient lives in the exchange buffer: symmetric memory + the library's kernels, NCCL as the fallback
    texture_img = torch.nn.Parameter(tex0.to(device))
    symm, bucket, mode = None, None, "none"
    if world > 1 and exchange != "nccl":
        try:
            symm = SymmetricGradientBuffer(texture_img.numel(), device)
            texture_img.grad = symm.view(tuple(texture_img.shape))
            mode = symm.mode
        except Exception as exc:                       # no symmetric memory on this box
            print(f"train_step.py: symmetric-memory exchange unavailable ({exc}); NCCL", file=sys.stderr)
            symm = None
    if world > 1 and symm is None:
        bucket, mode = GradientBucket([texture_img]), "nccl"
    opt = lp.optim.FusedAdam([texture_img], lr=5e-3, betas=(0.9, 0.99), eps=1e-15)
    guide = Guidance(device, unet_width)

    ev = {k: [torch.cuda.Event(enable_timing=True) for _ in range(2)] for k in ("render", "guide", "backward", "exchange", "adam")}
    acc = {k: 0.0 for k in ev}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for step in range(warmup + steps):
        if step == warmup:
            torch.cuda.synchronize(device)
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize(device)
            e0.record()
        timed = step >= warmup
        theta, phi, radius = views_for(step, rank, views, device)
        if symm is not None or bucket is not None:
            texture_img.grad.zero_()                              # optimizer.zero_grad(): the gradient buffer is persistent
        else:
            opt.zero_grad()
        ev["render"][0].record()
        pred_rgb, mask, normals, lighting = renderer.render_single_view_texture(
            verts, faces, uv, texture_img, theta, phi, radius, dims=(64, 64), is_body=bool(step % 2))       # trainer.py:621
        ev["render"][1].record(); ev["guide"][0].record()
        grad = guide.train_step(pred_rgb.detach())                                                           # trainer.py:657
        ev["guide"][1].record(); ev["backward"][0].record()
        if symm is not None or bucket is not None:
            # autograd accumulates into the persistent buffer (param.grad is a view of it)
            pred_rgb.backward(gradient=grad)                                                                 # trainer.py:658
        else:
            pred_rgb.backward(gradient=grad)
        ev["backward"][1].record(); ev["exchange"][0].record()
        if symm is not None:
            symm.all_reduce()
        elif bucket is not None:
            bucket.all_reduce()
        ev["exchange"][1].record(); ev["adam"][0].record()
        opt.step()                                                                                           # trainer.py:407
        ev["adam"][1].record()
        if timed:
            torch.cuda.synchronize(device)
            for k in ev:
                acc[k] += ev[k][0].elapsed_time(ev[k][1])
    e1.record()
    torch.cuda.synchronize(device)
    total_ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([total_ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    out = None
    if rank == 0:
        per = {k: v / steps for k, v in acc.items()}
        renderer_ms = per["render"] + per["backward"] + per["exchange"] + per["adam"]
        out = {"workload": "configs[4]: end-to-end SDS step, teddy.obj, 4ch %d^2 latent texture, 64x64 latent render, %d views/GPU, "
                           "random-init conditional UNet (width %d, bf16) in PyTorch" % (T, views, unet_width),
               "n_gpus": world, "steps": steps, "ms_per_step": total_ms / steps, "steps_per_s": 1e3 * steps / total_ms,
               "views_per_s": 1e3 * steps * views * world / total_ms, "exchange": mode,
               "ms": per, "renderer_ms": renderer_ms, "renderer_share_of_step": renderer_ms / (total_ms / steps),
               "texture_abs_mean_after": float(texture_img.detach().abs().mean()),
               "note": "timed per stage with a synchronise per step (the stage times add up to the step); the guidance "
                       "network is PyTorch and out of scope"}
        if not quiet:
            print(json.dumps(out), flush=True)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--views", type=int, default=8)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--unet-width", type=int, default=320)
    ap.add_argument("--texture", type=int, default=512)
    ap.add_argument("--exchange", default="auto", choices=["auto", "nccl"])
    a = ap.parse_args()
    run(a.views, a.steps, a.warmup, a.unet_width, a.texture, exchange=a.exchange)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
