"""Autograd-connected calls into the C ABI (``include/lp_b200.h``).

``render_texture`` stands where the reference chains ``prepare_vertices → rasterize /
dibr_rasterization → texture_mapping → mask composition`` (reference
``src/latent_paint/models/render.py:56-67``, ``src/latent_paint_mesh/models/render.py:194-277``);
``render_face_features`` where it rasterizes per-face-vertex colours (``render.py:39-45``).
Only the texture / the face features receive gradients, exactly as in the reference (UVs are
detached there, ``render.py:61``; vertices never require grad).

torch provides device memory and the stream; every kernel is ours.  No CPU fallback: a CPU
tensor for the texture or a missing library raises.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field

import torch

from . import _lib
from ._lib import LpBackwardArgs, LpForwardArgs

_INTERP = {"nearest": _lib.LP_INTERP_NEAREST, "bilinear": _lib.LP_INTERP_BILINEAR, "bicubic": _lib.LP_INTERP_BICUBIC}

#: kernel launches enqueued by this process through the wrappers below (bench.py reads it)
launch_counter = {"kernels": 0}

_workspaces: dict = {}
_int32_cache: dict = {}
_csr_cache: dict = {}


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _require_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"lp_b200: '{name}' must be a CUDA tensor — this renderer has no CPU path")


def _workspace(device, nbytes):
    """Scratch of the forward call, one buffer per (device, stream): two renders enqueued on two streams never share
    it, and a buffer that is outgrown stays referenced by the autograd node that used it (``keep`` lists below) until
    the caching allocator may hand it out again in stream order."""
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def _f32(t, device):
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def _faces_i32(faces, device, num_vertices=None):
    """The reference carries faces as int64 (LongTensor); the kernels read int32.  Converted once per source tensor:
    the cache entry keeps the source alive and is matched by identity and version, so a recycled address can never
    return another mesh's faces.  The index range is validated once, when the copy is made (a malformed mesh raises
    here where kaolin's gather would)."""
    if faces.dtype == torch.int32 and faces.device == device and faces.is_contiguous():
        return faces
    key = (id(faces), str(device))
    hit = _int32_cache.get(key)
    if hit is None or hit[0] is not faces or hit[1] != faces._version:
        if len(_int32_cache) > 16:
            _int32_cache.clear()
        conv = faces.detach().to(device=device, dtype=torch.int32).contiguous()
        if conv.numel():
            lo, hi = int(faces.min()), int(faces.max())
            if lo < 0 or (num_vertices is not None and hi >= num_vertices):
                raise ValueError(f"lp_b200: face indices span [{lo}, {hi}] but the mesh has {num_vertices} vertices")
        hit = (faces, faces._version, conv, int(faces.max()) if conv.numel() else -1)
        _int32_cache[key] = hit
    elif num_vertices is not None and hit[3] >= num_vertices:
        raise ValueError(f"lp_b200: face index {hit[3]} out of range for {num_vertices} vertices")
    return hit[2]


def vertex_face_csr(faces_i32, num_vertices):
    """Vertex → incident-corner lists in the order the reference accumulates them in
    ``compute_vertex_normals`` (render.py:99-101): corner 0 of every face in face order, then
    corner 1, then corner 2.  One stable sort per mesh topology, cached."""
    key = (id(faces_i32), num_vertices)
    hit = _csr_cache.get(key)
    if hit is not None and (hit[2] is not faces_i32 or hit[3] != faces_i32._version):
        hit = None
    if hit is None:
        if len(_csr_cache) > 16:
            _csr_cache.clear()
        F = faces_i32.shape[0]
        corner_major = faces_i32.t().reshape(-1).long()                    # (3F): k*F + f
        order = torch.sort(corner_major, stable=True).indices
        vf = (order % F).to(torch.int32).contiguous()
        counts = torch.bincount(corner_major, minlength=num_vertices)
        off = torch.zeros(num_vertices + 1, dtype=torch.int32, device=faces_i32.device)
        off[1:] = torch.cumsum(counts, 0).to(torch.int32)
        hit = (off.contiguous(), vf, faces_i32, faces_i32._version)
        _csr_cache[key] = hit
    return hit[0], hit[1]


@dataclass
class RenderConfig:
    """Everything of one render call that is not differentiable."""
    verts: torch.Tensor | None          # (V,3) f32 cuda        (None with the prepared-geometry input)
    faces: torch.Tensor | None          # (F,3) i32 cuda
    cameras: torch.Tensor | None        # (B,4,3) f32 cuda
    proj: tuple                         # (px, py, pz)
    H: int
    W: int
    flags: int
    interp: str = "nearest"
    face_uv: torch.Tensor | None = None  # (F,3,2) f32 cuda
    multiplier: float = 1000.0
    eps: float = 1e-8
    lights: torch.Tensor | None = None   # (9) f32 cuda → normals + lighting outputs
    want_buffers: bool = False           # face_idx / bary / depth extras
    extras: dict = field(default_factory=dict)
    # kaolin-level geometry input (kal.render.mesh.rasterize): already projected vertices
    fvi: torch.Tensor | None = None      # (B,F,3,2) f32 cuda
    fvz: torch.Tensor | None = None      # (B,F,3) f32 cuda
    valid: torch.Tensor | None = None    # (B,F) u8 cuda

    @property
    def B(self):
        return self.fvz.shape[0] if self.fvi is not None else self.cameras.shape[0]

    @property
    def F(self):
        return self.fvz.shape[1] if self.fvi is not None else self.faces.shape[0]


def _fill_common(a: LpForwardArgs, cfg: RenderConfig, device):
    B = cfg.B
    if cfg.fvi is not None:
        a.face_vertices_image, a.face_vertices_z, a.valid_faces = _ptr(cfg.fvi), _ptr(cfg.fvz), _ptr(cfg.valid)
        a.V, a.F = 0, cfg.F
    else:
        a.verts, a.faces = _ptr(cfg.verts), _ptr(cfg.faces)
        a.V, a.F = cfg.verts.shape[0], cfg.faces.shape[0]
        a.cameras = _ptr(cfg.cameras)
    a.B = B
    a.proj[0], a.proj[1], a.proj[2] = cfg.proj
    a.H, a.W = cfg.H, cfg.W
    a.multiplier, a.eps, a.flags = cfg.multiplier, cfg.eps, cfg.flags
    nbytes = _lib.lib().lp_workspace_bytes(B, a.F, cfg.H, cfg.W)
    ws = _workspace(device, nbytes)
    a.workspace, a.workspace_bytes = _ptr(ws), ws.numel()
    return ws


class _RenderTexture(torch.autograd.Function):
    @staticmethod
    def forward(ctx, texture, cfg: RenderConfig):
        _require_cuda(texture, "texture_map")
        device = texture.device
        if texture.dim() != 4 or texture.shape[0] != 1:
            raise ValueError(f"texture_map must have shape (1,C,T,T), got {tuple(texture.shape)}")
        if cfg.interp not in _INTERP:
            raise ValueError(f"lp_b200: interpolation mode '{cfg.interp}' is not implemented (nearest, bilinear, bicubic)")
        tex = texture.detach().to(torch.float32).contiguous()
        _, C, Th, Tw = tex.shape
        B, H, W = cfg.B, cfg.H, cfg.W
        image = torch.empty((B, C, H, W), dtype=torch.float32, device=device)
        mask = torch.empty((B, 1, H, W), dtype=torch.float32, device=device)
        uv = torch.empty((B, H, W, 2), dtype=torch.float32, device=device)
        face_idx = bary = depth = normals = lighting = None
        a = LpForwardArgs()
        with torch.cuda.device(device):
            ws = _fill_common(a, cfg, device)
            a.face_uv, a.texture = _ptr(cfg.face_uv), _ptr(tex)
            a.C, a.Th, a.Tw, a.interp = C, Th, Tw, _INTERP[cfg.interp]
            if cfg.want_buffers:
                face_idx = torch.empty((B, H, W), dtype=torch.int32, device=device)
                bary = torch.empty((B, H, W, 3), dtype=torch.float32, device=device)
                depth = torch.empty((B, H, W), dtype=torch.float32, device=device)
                a.face_idx, a.bary, a.depth = _ptr(face_idx), _ptr(bary), _ptr(depth)
            keep = [ws, tex]
            if cfg.lights is not None:
                off, vf = vertex_face_csr(cfg.faces, a.V)
                fn = torch.empty((B, a.F, 3), dtype=torch.float32, device=device)
                vn = torch.empty((B, a.V, 3), dtype=torch.float32, device=device)
                normals = torch.empty((B, 3, H, W), dtype=torch.float32, device=device)
                lighting = torch.empty((B, 1, H, W), dtype=torch.float32, device=device)
                a.vf_offsets, a.vf_faces, a.face_normals, a.vertex_normals = _ptr(off), _ptr(vf), _ptr(fn), _ptr(vn)
                a.lights, a.normals, a.lighting = _ptr(cfg.lights), _ptr(normals), _ptr(lighting)
                keep += [off, vf, fn, vn]
            a.image, a.mask, a.uv = _ptr(image), _ptr(mask), _ptr(uv)
            footprint_any = torch.empty((B, (H + 3) // 4, (W + 7) // 8), dtype=torch.uint8, device=device)
            a.footprint_any = _ptr(footprint_any)
            _lib.check(_lib.lib().lp_render_forward(ctypes.byref(a), _stream(device)))
            launch_counter["kernels"] += _lib.lib().lp_last_launch_count()
        ctx.cfg = cfg
        ctx.tex_shape = tuple(texture.shape)
        ctx.save_for_backward(uv, footprint_any)
        # the saved uv carries the kernels' own conventions (NaN on uncovered pixels of the masked flavour, tiles
        # without coverage not written at all); what the caller sees is kaolin's: 0 where nothing is covered
        uv_out = torch.where(face_idx[..., None] >= 0, uv, torch.zeros_like(uv)) if cfg.want_buffers else uv
        outs = (image, mask, uv_out, face_idx, bary, depth, normals, lighting)
        ctx.mark_non_differentiable(*[o for o in outs[1:] if o is not None])
        return outs

    @staticmethod
    def backward(ctx, grad_image, *unused):
        uv, footprint_any = ctx.saved_tensors
        cfg = ctx.cfg
        device = uv.device
        _, C, Th, Tw = ctx.tex_shape
        g = grad_image.to(torch.float32).contiguous()
        b = LpBackwardArgs()
        b.B, b.H, b.W, b.flags = uv.shape[0], cfg.H, cfg.W, cfg.flags
        wbytes = _lib.lib().lp_backward_workspace_bytes(C, Th, Tw)
        if wbytes:       # vector-RED path: the library zeroes its accumulation buffer and overwrites the gradient
            accum = torch.empty(int(wbytes), dtype=torch.uint8, device=device)
            b.workspace, b.workspace_bytes = _ptr(accum), accum.numel()
            b.flags |= _lib.LP_FLAG_GRAD_OVERWRITE
            grad_tex = torch.empty((1, C, Th, Tw), dtype=torch.float32, device=device)
        else:
            grad_tex = torch.zeros((1, C, Th, Tw), dtype=torch.float32, device=device)
        b.grad_image, b.uv = _ptr(g), _ptr(uv)
        b.C, b.Th, b.Tw, b.interp = C, Th, Tw, _INTERP[cfg.interp]
        b.grad_texture = _ptr(grad_tex)
        b.footprint_any = _ptr(footprint_any)
        with torch.cuda.device(device):
            _lib.check(_lib.lib().lp_render_backward(ctypes.byref(b), _stream(device)))
            launch_counter["kernels"] += _lib.lib().lp_last_launch_count()
        return grad_tex, None


class _RenderFeatures(torch.autograd.Function):
    @staticmethod
    def forward(ctx, face_features, cfg: RenderConfig):
        _require_cuda(face_features, "face_attributes")
        device = face_features.device
        ff = face_features.detach().to(torch.float32).contiguous()
        if ff.dim() != 4 or ff.shape[2] != 3:
            raise ValueError(f"face_attributes must have shape (1|B,F,3,D), got {tuple(ff.shape)}")
        B, H, W = cfg.B, cfg.H, cfg.W
        Bf, F, _, D = ff.shape
        if Bf not in (1, B) or F != cfg.F:
            raise ValueError("face_attributes batch/face count does not match the mesh / views")
        image = torch.empty((B, D, H, W), dtype=torch.float32, device=device)
        mask = torch.empty((B, 1, H, W), dtype=torch.float32, device=device)
        face_idx = torch.empty((B, H, W), dtype=torch.int32, device=device)
        bary = torch.empty((B, H, W, 3), dtype=torch.float32, device=device)
        depth = torch.empty((B, H, W), dtype=torch.float32, device=device) if cfg.want_buffers else None
        a = LpForwardArgs()
        with torch.cuda.device(device):
            ws = _fill_common(a, cfg, device)
            a.flags = cfg.flags | _lib.LP_FLAG_SHADE_FEATURES
            a.face_features, a.D, a.features_batched = _ptr(ff), D, int(Bf == B and B > 1)
            a.image, a.mask, a.face_idx, a.bary, a.depth = _ptr(image), _ptr(mask), _ptr(face_idx), _ptr(bary), _ptr(depth)
            _lib.check(_lib.lib().lp_render_forward(ctypes.byref(a), _stream(device)))
            launch_counter["kernels"] += _lib.lib().lp_last_launch_count()
        ctx.cfg, ctx.ff_shape, ctx.batched = cfg, tuple(ff.shape), int(Bf == B and B > 1)
        ctx.save_for_backward(face_idx, bary)
        ctx.mark_non_differentiable(*[o for o in (mask, face_idx, bary, depth) if o is not None])
        return image, mask, face_idx, bary, depth

    @staticmethod
    def backward(ctx, grad_image, *unused):
        face_idx, bary = ctx.saved_tensors
        cfg = ctx.cfg
        device = face_idx.device
        Bf, F, _, D = ctx.ff_shape
        g = grad_image.to(torch.float32).contiguous()
        grad_ff = torch.zeros(ctx.ff_shape, dtype=torch.float32, device=device)
        b = LpBackwardArgs()
        b.B, b.H, b.W = face_idx.shape[0], cfg.H, cfg.W
        b.flags = cfg.flags | _lib.LP_FLAG_SHADE_FEATURES
        b.grad_image, b.face_idx, b.bary = _ptr(g), _ptr(face_idx), _ptr(bary)
        b.F, b.D, b.features_batched = F, D, ctx.batched
        b.grad_face_features = _ptr(grad_ff)
        with torch.cuda.device(device):
            _lib.check(_lib.lib().lp_render_backward(ctypes.byref(b), _stream(device)))
            launch_counter["kernels"] += _lib.lib().lp_last_launch_count()
        return grad_ff, None


class _RenderComposed(torch.autograd.Function):
    """Texture render of the object + face-colour render of the environment sphere + the model-level composition
    ``pred_back * (1 - mask) + pred_features * mask`` (reference ``src/latent_paint/models/textured_mesh.py:195-212``)
    as two forward launches chains and no elementwise pass: the composition is written by the shading stage of the
    second render, and the backward feeds ``dL/d composed`` straight into both scatter kernels (the texture scatter
    already ignores uncovered pixels, the face-colour scatter scales by ``1 - mask`` per pixel)."""

    @staticmethod
    def forward(ctx, texture, face_features, tex_cfg: RenderConfig, feat_cfg: RenderConfig):
        ctx.set_materialize_grads(False)
        tex_ctx = _Scratch()
        with torch.no_grad():
            fg, mask, uv, *_ = _RenderTexture.forward(tex_ctx, texture, tex_cfg)
        _require_cuda(face_features, "face_attributes")
        device = face_features.device
        ff = face_features.detach().to(torch.float32).contiguous()
        B, H, W = feat_cfg.B, feat_cfg.H, feat_cfg.W
        Bf, F, _, D = ff.shape
        if ff.dim() != 4 or ff.shape[2] != 3 or Bf not in (1, B) or F != feat_cfg.F:
            raise ValueError(f"face_attributes must have shape (1|B,F,3,D) matching the sphere mesh, got {tuple(ff.shape)}")
        if tuple(fg.shape) != (B, D, H, W):
            raise ValueError(f"texture render {tuple(fg.shape)} and face-colour render {(B, D, H, W)} do not match")
        back = torch.empty((B, D, H, W), dtype=torch.float32, device=device)
        bmask = torch.empty((B, 1, H, W), dtype=torch.float32, device=device)
        composed = torch.empty((B, D, H, W), dtype=torch.float32, device=device)
        face_idx = torch.empty((B, H, W), dtype=torch.int32, device=device)
        bary = torch.empty((B, H, W, 3), dtype=torch.float32, device=device)
        a = LpForwardArgs()
        with torch.cuda.device(device):
            ws = _fill_common(a, feat_cfg, device)
            a.flags = feat_cfg.flags | _lib.LP_FLAG_SHADE_FEATURES
            a.face_features, a.D, a.features_batched = _ptr(ff), D, int(Bf == B and B > 1)
            a.image, a.mask, a.face_idx, a.bary = _ptr(back), _ptr(bmask), _ptr(face_idx), _ptr(bary)
            a.under_image, a.under_mask, a.composed = _ptr(fg), _ptr(mask), _ptr(composed)
            _lib.check(_lib.lib().lp_render_forward(ctypes.byref(a), _stream(device)))
            launch_counter["kernels"] += _lib.lib().lp_last_launch_count()
        ctx.tex_ctx, ctx.feat_cfg = tex_ctx, feat_cfg
        ctx.ff_shape, ctx.batched = tuple(ff.shape), int(Bf == B and B > 1)
        ctx.save_for_backward(face_idx, bary, mask, *tex_ctx.saved_tensors)
        ctx.mark_non_differentiable(mask)
        return composed, mask, back, fg

    @staticmethod
    def backward(ctx, g_composed, g_mask, g_back, g_fg):
        face_idx, bary, mask, uv, footprint_any = ctx.saved_tensors
        device = face_idx.device
        grad_tex = grad_ff = None
        # texture: d composed / d foreground = mask, and the texture scatter only sees covered pixels (mask = 1)
        g_t = g_composed if g_fg is None else (g_fg if g_composed is None else g_composed + g_fg)
        if g_t is not None and ctx.needs_input_grad[0]:
            ctx.tex_ctx.saved_tensors = (uv, footprint_any)
            grad_tex, _ = _RenderTexture.backward(ctx.tex_ctx, g_t)
        if (g_composed is not None or g_back is not None) and ctx.needs_input_grad[1]:
            cfg = ctx.feat_cfg
            Bf, F, _, D = ctx.ff_shape
            b = LpBackwardArgs()
            if g_back is None:        # the common case: the (1 - mask) factor is applied inside the scatter kernel
                g = g_composed.to(torch.float32).contiguous()
                b.under_mask = _ptr(mask)
            else:
                g = (g_back if g_composed is None else g_composed * (1 - mask) + g_back).to(torch.float32).contiguous()
            grad_ff = torch.zeros(ctx.ff_shape, dtype=torch.float32, device=device)
            b.B, b.H, b.W = face_idx.shape[0], cfg.H, cfg.W
            b.flags = cfg.flags | _lib.LP_FLAG_SHADE_FEATURES
            b.grad_image, b.face_idx, b.bary = _ptr(g), _ptr(face_idx), _ptr(bary)
            b.F, b.D, b.features_batched = F, D, ctx.batched
            b.grad_face_features = _ptr(grad_ff)
            with torch.cuda.device(device):
                _lib.check(_lib.lib().lp_render_backward(ctypes.byref(b), _stream(device)))
                launch_counter["kernels"] += _lib.lib().lp_last_launch_count()
        return grad_tex, grad_ff, None, None


class _Scratch:
    """Stand-in for an autograd ctx so ``_RenderTexture.forward`` / ``.backward`` can be reused inside another Function."""

    def __init__(self):
        self.saved_tensors = ()

    def save_for_backward(self, *tensors):
        self.saved_tensors = tensors

    def mark_non_differentiable(self, *tensors):
        pass


def render_composed(texture, face_features, tex_cfg: RenderConfig, feat_cfg: RenderConfig):
    """→ (composed (B,C,H,W), mask (B,1,H,W), background (B,C,H,W), foreground (B,C,H,W)); gradients flow into
    ``texture`` and ``face_features``."""
    return _RenderComposed.apply(texture, face_features, tex_cfg, feat_cfg)


def render_texture(texture, cfg: RenderConfig):
    """→ (image (B,C,H,W), mask (B,1,H,W), uv (B,H,W,2), face_idx|None, bary|None, depth|None,
    normals|None, lighting|None).  ``uv`` is meaningful only with ``cfg.want_buffers`` (then 0 on uncovered
    pixels, like kaolin's interpolated features); otherwise it is the kernels' saved-for-backward buffer."""
    return _RenderTexture.apply(texture, cfg)


def render_face_features(face_features, cfg: RenderConfig):
    """→ (image (B,D,H,W), mask (B,1,H,W), face_idx (B,H,W) i32, bary (B,H,W,3), depth|None)."""
    return _RenderFeatures.apply(face_features, cfg)


def cameras_from_views(elev, azim, radius, look_at_height: float):
    """Device-side ``get_camera_from_view`` for angle tensors that already live on the GPU."""
    device = elev.device
    _require_cuda(elev, "elev")
    B = elev.numel()
    e = elev.detach().to(torch.float32).contiguous().reshape(-1)
    a = azim.detach().to(device=device, dtype=torch.float32).contiguous().reshape(-1)
    if torch.is_tensor(radius):
        r = radius.detach().to(device=device, dtype=torch.float32).contiguous().reshape(-1)
    else:
        r = torch.full((1,), float(radius), dtype=torch.float32, device=device)
    stride = 1 if r.numel() == B and B > 1 else 0
    if r.numel() not in (1, B):
        raise ValueError("radius must be a scalar or have one entry per view")
    out = torch.empty((B, 4, 3), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        _lib.check(_lib.lib().lp_cameras_from_views(_ptr(e), _ptr(a), _ptr(r), stride, float(look_at_height), B,
                                                    _ptr(out), _stream(device)))
        launch_counter["kernels"] += 1
    return out


class _TextureMap(torch.autograd.Function):
    """``kal.render.mesh.texture_mapping``: uv (B,H,W,2), textures (B|1,C,T,T) → (B,C,H,W)."""

    @staticmethod
    def forward(ctx, texture_maps, uv, mode):
        _require_cuda(texture_maps, "texture_maps")
        if mode not in _INTERP:
            raise ValueError(f"lp_b200: interpolation mode '{mode}' is not implemented (nearest, bilinear, bicubic)")
        device = texture_maps.device
        tex = texture_maps.detach().to(torch.float32).contiguous()
        uvc = uv.detach().to(device=device, dtype=torch.float32).contiguous()
        B, H, W = uvc.shape[0], uvc.shape[1], uvc.shape[2]
        Bt, C, Th, Tw = tex.shape
        if Bt not in (1, B):
            raise ValueError("texture_maps batch must be 1 or match the coordinates")
        out = torch.empty((B, C, H, W), dtype=torch.float32, device=device)
        a = _lib.LpTextureMapArgs()
        a.B, a.H, a.W, a.uv, a.texture = B, H, W, _ptr(uvc), _ptr(tex)
        a.texture_batch_stride = C * Th * Tw if Bt == B and B > 1 else 0
        a.C, a.Th, a.Tw, a.interp, a.out = C, Th, Tw, _INTERP[mode], _ptr(out)
        with torch.cuda.device(device):
            _lib.check(_lib.lib().lp_texture_map_forward(ctypes.byref(a), _stream(device)))
            launch_counter["kernels"] += 1
        ctx.mode, ctx.tex_shape = mode, tuple(tex.shape)
        ctx.save_for_backward(uvc)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (uvc,) = ctx.saved_tensors
        device = uvc.device
        Bt, C, Th, Tw = ctx.tex_shape
        g = grad_out.to(torch.float32).contiguous()
        grad_tex = torch.zeros(ctx.tex_shape, dtype=torch.float32, device=device)
        b = LpBackwardArgs()
        b.B, b.H, b.W, b.flags = uvc.shape[0], uvc.shape[1], uvc.shape[2], 0
        b.grad_image, b.uv = _ptr(g), _ptr(uvc)
        b.C, b.Th, b.Tw, b.interp = C, Th, Tw, _INTERP[ctx.mode]
        b.grad_texture = _ptr(grad_tex)
        b.grad_texture_batch_stride = C * Th * Tw if Bt > 1 else 0
        with torch.cuda.device(device):
            _lib.check(_lib.lib().lp_render_backward(ctypes.byref(b), _stream(device)))
            launch_counter["kernels"] += _lib.lib().lp_last_launch_count()
        return grad_tex, None, None


def texture_map(uv, texture_maps, mode="nearest"):
    return _TextureMap.apply(texture_maps, uv, mode)


class _ResizeBicubic(torch.autograd.Function):
    """``F.interpolate(x, size, mode='bicubic')`` (align_corners=False) of several equally sized (B,C_i,H,W) tensors in
    one launch, forward and backward (``lp_resize_bicubic``)."""

    @staticmethod
    def _launch(ins, outs, H, W, OH, OW, backward, device):
        a = _lib.LpResizeArgs()
        a.n, a.H, a.W, a.OH, a.OW, a.backward = len(ins), H, W, OH, OW, int(backward)
        for i, (x, y) in enumerate(zip(ins, outs)):
            a.inp[i], a.out[i], a.planes[i] = x.data_ptr(), y.data_ptr(), x.shape[0] * x.shape[1]
        with torch.cuda.device(device):
            _lib.check(_lib.lib().lp_resize_bicubic(ctypes.byref(a), _stream(device)))
            launch_counter["kernels"] += 1

    @staticmethod
    def forward(ctx, size, *tensors):
        device = tensors[0].device
        _require_cuda(tensors[0], "resize input")
        xs = [t.detach().to(torch.float32).contiguous() for t in tensors]
        H, W = xs[0].shape[-2:]
        if any(x.dim() != 4 or tuple(x.shape[-2:]) != (H, W) for x in xs) or len(xs) > 8:
            raise ValueError("resize_bicubic: up to eight (B,C,H,W) tensors of one spatial size")
        OH, OW = size
        outs = [torch.empty(x.shape[0], x.shape[1], OH, OW, dtype=torch.float32, device=device) for x in xs]
        _ResizeBicubic._launch(xs, outs, H, W, OH, OW, False, device)
        ctx.shape = (H, W, OH, OW)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        H, W, OH, OW = ctx.shape
        idx = [i for i, g in enumerate(grads) if g is not None and ctx.needs_input_grad[i + 1]]
        res = [None] * len(grads)
        if idx:
            device = grads[idx[0]].device
            gs = [grads[i].to(torch.float32).contiguous() for i in idx]
            outs = [torch.zeros(g.shape[0], g.shape[1], H, W, dtype=torch.float32, device=device) for g in gs]
            _ResizeBicubic._launch(gs, outs, H, W, OH, OW, True, device)
            for i, o in zip(idx, outs):
                res[i] = o
        return (None, *res)


def resize_bicubic(tensors, size):
    """[(B,C_i,H,W), ...] -> [(B,C_i,size[0],size[1]), ...]: torch's bicubic ``F.interpolate`` for all of them in one launch."""
    return list(_ResizeBicubic.apply(tuple(int(s) for s in size), *tensors))


def depth_for_guidance(depth, size=64, normalised=True):
    """The depth input of depth-conditioned guidance (reference ``src/stable_diffusion_depth.py:302-319``:
    ``train_step(text, inputs, depth_mask)`` resizes ``depth_mask`` to 64 x 64 with bicubic ``F.interpolate`` and min-max
    normalises the whole tensor to [-1, 1]) from the rasterizer's depth buffer.  ``depth`` (B,H,W): camera-space z of
    the visible surface (< 0), 0 where nothing is covered.  The map is MiDaS-like inverse distance, ``1 / -z`` on the
    surface and 0 on the background (nearer = larger), resized by ``lp_resize_bicubic`` and normalised.
    -> (B,1,size,size)."""
    d = depth.to(torch.float32)
    inv = torch.where(d < 0, -1.0 / d.clamp(max=-1e-12), torch.zeros_like(d))[:, None].contiguous()
    if inv.shape[-1] != size or inv.shape[-2] != size:
        inv = resize_bicubic([inv], (size, size))[0]
    if normalised:
        lo, hi = inv.min(), inv.max()
        inv = 2.0 * (inv - lo) / (hi - lo) - 1.0
    return inv
