"""Shared helpers for the parity tests: seeded scenes in the reference's own distributions
 and the tolerance BASELINE.json states for floating point."""
import os

import numpy as np
import torch

import latent_nerf_test_b200 as lp

RTOL, ATOL = 1e-4, 1e-5          # north_star: rendered pixels and texture gradients
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def scene(shape, scale, dy, subdivide=0):
    m = lp.meshio.find_shape(shape)
    if subdivide:
        m = lp.meshio.subdivide(m, subdivide)
    verts = lp.meshio.normalize_vertices(m.vertices, scale, dy)
    return verts, m.faces, lp.meshio.face_uv_attributes(m)


def scene_raw(shape):
    """Un-normalised mesh (the environment sphere is used at its own radius-20 scale, reference textured_mesh.py:53-54)."""
    m = lp.meshio.find_shape(shape)
    return m.vertices, m.faces, m


def rnd(shape, seed, scale=1.0):
    return scale * torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def latent_paint_views(B, seed=0):
    """radius, theta, phi drawn in the reference's order (latent_paint views_dataset.py:16-18);
    theta from 15° to avoid the up‖view singularity."""
    g = torch.Generator().manual_seed(seed)
    radius = torch.rand(B, generator=g) * 0.5 + 1.0
    theta = torch.deg2rad(torch.rand(B, generator=g) * 135.0 + 15.0)
    phi = torch.deg2rad(torch.rand(B, generator=g) * 360.0)
    return radius, theta, phi


def mesh_views(B, seed=0):
    """latent_paint_mesh train_config.py:18-22 ranges."""
    g = torch.Generator().manual_seed(seed)
    radius = torch.rand(B, generator=g) * 1.0 + 1.4
    theta = torch.deg2rad(torch.rand(B, generator=g) * 50.0 + 60.0)
    phi = torch.deg2rad(torch.rand(B, generator=g) * 360.0)
    return radius, theta, phi


def assert_close(a, b, what, rtol=RTOL, atol=ATOL):
    a = torch.as_tensor(a).detach().cpu().float()
    b = torch.as_tensor(b).detach().cpu().float()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    bad = err > tol
    assert not bad.any(), f"{what}: {int(bad.sum())} of {bad.numel()} elements off, max abs err {float(err.max()):.3e}"


U32 = 2.0 ** -24                  # unit roundoff of fp32


def accumulation_tolerance(n_terms, rms, atol=ATOL):
    """Tolerance for ONE texel that is the fp32 sum of ``n_terms`` contributions of root-mean-square magnitude ``rms``
    added one at a time in an arbitrary order (atomics on the GPU, a loop on the CPU) — the mesh flavour's background
    texel, which every uncovered pixel of every view feeds with weight 1 ("background-texel leak").
    Model: the partial sum after k additions is a random walk of size rms * sqrt(k); addition k rounds by at most
    u * |partial sum|, so the accumulated rounding is itself a random walk with standard deviation
    u * rms * sqrt(sum_k k) = u * rms * n / sqrt(2).  The bound is six standard deviations plus the pixel atol.  (The
    worst-case bound (n - 1) * u * sum|x_i| is orders of magnitude looser and would hide real errors.)"""
    return atol + 6.0 * U32 * rms * n_terms / 2 ** 0.5


def assert_texture_grad_close(got, ref, what, background=None, terms=None, rtol=RTOL, atol=ATOL):
    """``got`` / ``ref`` (..., C, T, T) texture gradients.  Everything at rtol / atol; ``background`` =
    ``(n_pixels, rms, ref64)`` relaxes ONLY texel (row T-1, col 0) to the accumulation bound around the fp64
    sum ``ref64`` (C,), and checks that the fp32 CPU reference obeys the same bound (so the bound is not hiding a bug).
    ``terms`` = ``(S_abs, n)`` from :func:`accumulation_terms` adds, per texel, the rounding a sum of n fp32 terms of total
    magnitude S_abs can pick up in an arbitrary summation order, ``u * sqrt(n) * S_abs`` (texels at a UV pole or under a
    minified surface sum hundreds of contributions; an ordinary texel sums a handful and keeps the plain tolerance)."""
    got = torch.as_tensor(got).detach().cpu().float().reshape(-1, *torch.as_tensor(ref).shape[-3:])
    ref = torch.as_tensor(ref).detach().cpu().float().reshape(got.shape)
    if background is not None:
        n, rms, ref64 = background
        tol = accumulation_tolerance(n, rms)
        g_bg, r_bg = got[:, :, -1, 0].double().sum(0), ref[:, :, -1, 0].double().sum(0)
        ref64 = torch.as_tensor(ref64).double().reshape(-1)
        assert float((g_bg - ref64).abs().max()) <= tol, \
            f"{what}: background texel off by {float((g_bg - ref64).abs().max()):.3e} > derived bound {tol:.3e} ({n} terms, rms {rms:.2f})"
        assert float((r_bg - ref64).abs().max()) <= tol, f"{what}: the fp32 CPU reference itself violates the derived bound"
        got, ref = got.clone(), ref.clone()
        got[:, :, -1, 0] = 0
        ref[:, :, -1, 0] = 0
    if terms is None:
        return assert_close(got, ref, what, rtol=rtol, atol=atol)
    S_abs, n = (torch.as_tensor(t).float().reshape(got.shape) for t in terms)
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs() + U32 * torch.sqrt(n.clamp(min=1)) * S_abs
    bad = err > tol
    assert not bad.any(), f"{what}: {int(bad.sum())} of {bad.numel()} elements off, max abs err {float(err.max()):.3e}"


def accumulation_terms(uv, grad, tex_shape, mode="bilinear"):
    """Per texel: the sum of |weight * upstream gradient| over everything the backward adds into it, and the number of
    contributions — from the real ATen grid_sample backward fed |grad| (bilinear weights are non-negative) and ones.
    ``uv`` (B,H,W,2), ``grad`` (B,C,H,W) -> two (1,C,T,T) tensors."""
    from oracle import kaolin_shim as kal
    uv, grad = torch.as_tensor(uv).detach().cpu().double(), torch.as_tensor(grad).detach().cpu().double()
    B = uv.shape[0]
    out = []
    for g in (grad.abs(), torch.ones_like(grad)):
        acc = torch.zeros(tex_shape[-3:], dtype=torch.float64)
        for i in range(0, B, 8):
            t = torch.zeros((min(8, B - i),) + tuple(tex_shape[-3:]), dtype=torch.float64, requires_grad=True)
            kal.texture_mapping(uv[i:i + 8], t, mode).backward(g[i:i + 8].permute(0, 2, 3, 1))
            acc += t.grad.sum(0)
        out.append(acc[None])
    weights_present = out[1] > 0
    return out[0], torch.where(weights_present, (out[1] * 4).clamp(min=1), torch.ones_like(out[1]))     # <= 4 taps per unit of weight


def fp64_corner_texel(uv, tex, grad, face_idx):
    """The gradient of texel (T-1, 0) summed in fp64 (same fp32 inputs, the real ATen grid_sample arithmetic in double),
    plus the number of uncovered pixels feeding it and the rms of their upstream gradient."""
    from oracle import kaolin_shim as kal
    uv, tex, grad, face_idx = (torch.as_tensor(x).detach().cpu() for x in (uv, tex, grad, face_idx))
    B = uv.shape[0]
    out = torch.zeros(tex.shape[1], dtype=torch.float64)
    for i in range(0, B, 8):
        t64 = tex.detach().double().repeat(min(8, B - i), 1, 1, 1).requires_grad_(True)
        img = kal.texture_mapping(uv[i:i + 8].double(), t64, "bilinear")          # (b,H,W,C)
        img.backward(grad[i:i + 8].permute(0, 2, 3, 1).double())
        out += t64.grad[:, :, -1, 0].sum(0)
    bg = face_idx < 0
    n = int(bg.sum())
    gb = grad.permute(0, 2, 3, 1)[bg]
    return n, float(torch.sqrt((gb.double() ** 2).mean())) if n else 0.0, out
