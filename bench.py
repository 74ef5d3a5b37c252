#!/usr/bin/env python
"""bench.py — views/sec of the Latent-Paint render path (forward + backward), BASELINE.json's metric.

Default workload (``config.workload``): BASELINE.json configs[1] — shapes/nascar.obj (V=3750, F=7500,
deterministic grid-atlas UVs), 3-channel 1024² RGB texture, 512×512 render, batch of
8 random views per GPU, latent_paint flavour, bilinear.  One *step* = one forward render of the
views + one backward scatter of a dense upstream gradient into the texture gradient (+ one sum of
that gradient over the ranks when N > 1 — the library's exchange kernels over symmetric memory, NCCL
as the fallback: views shard over ranks, weak scaling).  ``--workload c3`` (configs[2]: teddy, 64 views
in total sharded over the ranks, latent_paint_mesh flavour — strong scaling) and ``--workload c4``
(configs[3]: 1.31 M-face sphere, 1024², 16 views per GPU) run the same step on the other configs;
the default run also appends the configs[2] strong-scaling measurement as ``strong_scaling``.

  value      device-timed views/s, inputs resident in HBM, steps replayed from CUDA graphs (the graph is
             uploaded and replayed once during warm-up); by default a three-stream pipeline (geometry |
             visibility | texture fetch + backward + exchange) overlaps the texture-independent stages
             of later steps (--pipeline off: serial).  A step count that is not a multiple of the replay
             length (up to 40 steps) replays a second, shorter graph for the remainder.  At N > 1 the exchange
             is one kernel launch on the main stream (handshakes inside); with an exchange window long enough
             to hold it, the visibility stage of step i + 1 is issued one step ahead and released by the
             backward of step i - 1, so that it runs while exchange i - 1 is on the wire
  e2e        same metric through lp_render_step_host with pinned HOST buffers (H2D + D2H inside), on
             every rank at once, aggregated
  roofline   dominant kernel: algorithmic bytes per launch / its mean CUDA-event duration (an
             instrumented eager pass over the same steps) against MEASURED_PEAKS.json HBM GB/s
  cpu_baseline  the oracle (reference glue mirror over the torch kaolin restatement) on the host cores

``--impl reference`` times that CPU path alone and prints the same line with "impl": "reference".
``--workload c5`` is BASELINE.json configs[4] (tools/train_step.py: renderer + stand-in UNet + exchange + fused Adam).
Experiment switches (environment): LP_RASTER_CTAS, LP_WALK_CTAS, LP_EXCHANGE_CTAS, LP_EXCHANGE_BULK, LP_PDL,
LP_GATE_RASTER, LP_RASTER_LOOKAHEAD, LP_PREP_AFTER_BWD, LP_CLEAR_ASIDE, LP_MICRO, LP_PIPE_STEPS, LP_B200_LIB (another build of the
library, e.g. the -DLP_CHECKED one: its violation counter is reported), LP_DEBUG_CHECK.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import latent_nerf_test_b200 as lp  # noqa: E402
from latent_nerf_test_b200 import _lib  # noqa: E402

WORKLOADS = {
    # B = views per GPU (weak scaling) or views in total (strong scaling: sharded over the ranks)
    "c2": dict(shape="nascar", scale=0.6, dy=0.25, H=512, W=512, C=3, T=1024, B=8, interp="bilinear", flavour="lp",
               scaling="weak",
               label="configs[1]: nascar.obj F=7500, 3ch 1024^2 texture, 512x512, 8 views/GPU, latent_paint flavour, bilinear"),
    "c1": dict(shape="blub", scale=0.6, dy=0.25, H=64, W=64, C=4, T=128, B=1, interp="nearest", flavour="lp",
               scaling="weak",
               label="configs[0]: blub.obj F=14208, 4ch 128^2 latent texture, 64x64, 1 view, nearest"),
    "c3": dict(shape="teddy", scale=1.0, dy=0.0, H=64, W=64, C=4, T=512, B=64, interp="bilinear", flavour="mesh",
               scaling="strong",
               label="configs[2]: teddy.obj F=5760, 4ch 512^2 latent texture, 64x64 (train_grid_size), 64 views in total "
                     "sharded over the GPUs, latent_paint_mesh flavour (body camera), bilinear"),
    "c4": dict(shape="sphere", subdivide=5, scale=0.6, dy=0.25, H=1024, W=1024, C=3, T=1024, B=16, interp="bilinear",
               flavour="lp", scaling="weak",
               label="configs[3]: sphere.obj subdivided 5x F=1310720, 3ch 1024^2 texture, 1024x1024, 16 views/GPU, "
                     "latent_paint flavour, bilinear"),
}
# configs[4] (the end-to-end SDS step) is driven by tools/train_step.py; ``--workload c5`` runs it and prints its line
C5_LABEL = "configs[4]: end-to-end train_latent_paint_mesh SDS step, renderer in the loop, random-init UNet in PyTorch"
FOV = np.pi / 3                       # latent_paint flavour (render.py:11)
FOV_BODY, LOOK_AT_BODY = np.pi / 4, -0.3   # latent_paint_mesh flavour, body camera (render.py:18-19, 33)


def algorithmic_bytes(V, F, H, W, C, T, B, flavour="lp"):
    """SURVEY.md §8(d) / BASELINE.md §3, split per kernel (DESIGN.md 'Algorithmic bytes'); the mesh flavour adds
    normals (12) + lighting (4) bytes per pixel to the forward."""
    fwd = B * (12 * V + 36 * F + H * W * (4 * C + 12 + (16 if flavour == "mesh" else 0))) + 4 * C * T * T
    bwd = B * (H * W * (4 * C + 8)) + 4 * C * T * T
    return fwd, bwd


def make_views(B, seed, flavour="lp"):
    """radius, theta, phi drawn in the reference's order (views_dataset.py:16-18 / :22-24)."""
    g = torch.Generator().manual_seed(seed)
    if flavour == "mesh":                         # latent_paint_mesh train_config.py:18-22
        radius = torch.rand(B, generator=g) * 1.0 + 1.4
        theta = torch.deg2rad(torch.rand(B, generator=g) * 50.0 + 60.0)
    else:                                         # latent_paint views_dataset.py:9 (theta from 15 deg: SURVEY.md 8d)
        radius = torch.rand(B, generator=g) * 0.5 + 1.0
        theta = torch.deg2rad(torch.rand(B, generator=g) * 135.0 + 15.0)
    phi = torch.deg2rad(torch.rand(B, generator=g) * 360.0)
    return radius, theta, phi


def cameras_for(radius, theta, phi, dy):
    return torch.cat([lp.camera.camera_from_view(theta[i], phi[i], float(radius[i]), dy) for i in range(len(theta))]).contiguous()


def workload_cameras(w, B, seed):
    radius, theta, phi = make_views(B, seed, w.get("flavour", "lp"))
    return cameras_for(radius, theta, phi, LOOK_AT_BODY if w.get("flavour") == "mesh" else w["dy"])


def load_scene(w):
    m = lp.meshio.find_shape(w["shape"])
    if w.get("subdivide"):
        m = lp.meshio.subdivide(m, w["subdivide"])
    verts = lp.meshio.normalize_vertices(m.vertices, w["scale"], w["dy"])
    return verts, m.faces, lp.meshio.face_uv_attributes(m)


def render_forward_raster_fused(fwd, stream):
    """Tile rasterizer with the texture fetch fused (lp_render_forward minus lp_render_prepare)."""
    return _lib.lib().lp_render_raster_shade(ctypes.byref(fwd), stream)


class DeviceStep:
    """One buffer set + the two C-ABI argument blocks of a fwd+bwd step on resident inputs.
    ``w['flavour']``: 'lp' = latent_paint Renderer (masked image, fov pi/3), 'mesh' = latent_paint_mesh Renderer
    (unmasked image, float mask, normals + SH lighting outputs, nz == 0 culling, body camera)."""

    def __init__(self, geom, w, cams, seed, device, grad_tex=None, accum=None, grad_layout="planar"):
        verts, faces, uv = geom
        B, H, W, C, T = w["B"], w["H"], w["W"], w["C"], w["T"]
        mesh_flavour = w.get("flavour", "lp") == "mesh"
        self.device = device
        self.tex = (0.4 * torch.randn(1, C, T, T, generator=torch.Generator().manual_seed(seed))).to(device)
        self.grad_image = torch.randn(B, C, H, W, generator=torch.Generator().manual_seed(seed + 1)).to(device)
        self.cams = cams.to(device)
        self.image = torch.empty(B, C, H, W, device=device)
        self.mask = torch.empty(B, 1, H, W, device=device)
        self.uv = torch.empty(B, H, W, 2, device=device)
        self.grad_tex = torch.zeros(C, T, T, device=device) if grad_tex is None else grad_tex
        self.footprint_any = torch.empty(B, (H + 3) // 4, (W + 7) // 8, dtype=torch.uint8, device=device)
        L = _lib.lib()
        self.ws = torch.empty(int(L.lp_workspace_bytes(B, faces.shape[0], H, W)), dtype=torch.uint8, device=device)
        a = _lib.LpForwardArgs()
        a.verts, a.faces, a.V, a.F = verts.data_ptr(), faces.data_ptr(), verts.shape[0], faces.shape[0]
        a.cameras, a.B, a.H, a.W = self.cams.data_ptr(), B, H, W
        p = 1.0 / np.tan((FOV_BODY if mesh_flavour else FOV) / 2)
        a.proj[0], a.proj[1], a.proj[2] = p, p, -1.0
        a.multiplier, a.eps = 1000.0, 1e-8
        a.flags = (_lib.LP_FLAG_CULL_NZ_ZERO if mesh_flavour else _lib.LP_FLAG_MASK_IMAGE) | _lib.LP_FLAG_REJECT_BEHIND
        a.face_uv, a.texture = uv.data_ptr(), self.tex.data_ptr()
        a.C, a.Th, a.Tw = C, T, T
        # the texture once more, texel-interleaved, for the 16-byte taps of lp_render_shade (repacked by repack()
        # whenever the texture changes; in a training loop the optimiser step is followed by one repack)
        self.tex_rgba = torch.empty(T * T, 4, device=device) if C <= 4 and os.environ.get("LP_TEX_RGBA", "1") == "1" else None
        if self.tex_rgba is not None:
            a.texture_rgba = self.tex_rgba.data_ptr()
        a.interp = _lib.LP_INTERP_BILINEAR if w["interp"] == "bilinear" else _lib.LP_INTERP_NEAREST
        a.image, a.mask, a.uv = self.image.data_ptr(), self.mask.data_ptr(), self.uv.data_ptr()
        a.workspace, a.workspace_bytes = self.ws.data_ptr(), self.ws.numel()
        a.footprint_any = self.footprint_any.data_ptr()
        self.keep = [verts, faces, uv]
        if mesh_flavour:
            V, F = verts.shape[0], faces.shape[0]
            off, vf = lp.functional.vertex_face_csr(faces, V)
            self.fn = torch.empty(B, F, 3, device=device)
            self.vn = torch.empty(B, V, 3, device=device)
            self.normals = torch.empty(B, 3, H, W, device=device)
            self.lighting = torch.empty(B, 1, H, W, device=device)
            self.lights = torch.tensor([1.0, 0.0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0], device=device)
            a.vf_offsets, a.vf_faces = off.data_ptr(), vf.data_ptr()
            a.face_normals, a.vertex_normals = self.fn.data_ptr(), self.vn.data_ptr()
            a.lights, a.normals, a.lighting = self.lights.data_ptr(), self.normals.data_ptr(), self.lighting.data_ptr()
            self.keep += [off, vf]
        if os.environ.get("LP_MICRO") in ("0", "1"):     # experiments: force the micro-face path off / on (default: the density rule)
            a.flags |= _lib.LP_FLAG_MICRO_ON if os.environ["LP_MICRO"] == "1" else _lib.LP_FLAG_MICRO_OFF
        if os.environ.get("LP_DEBUG_FWD_STOP"):      # only honoured by -DLP_PROFILE builds (tools/)
            a.flags |= 1 << int(os.environ["LP_DEBUG_FWD_STOP"])
        b = _lib.LpBackwardArgs()
        b.B, b.H, b.W, b.flags = B, H, W, a.flags & 0xff
        b.grad_image, b.uv = self.grad_image.data_ptr(), self.uv.data_ptr()
        b.C, b.Th, b.Tw, b.interp = C, T, T, a.interp
        b.grad_texture = self.grad_tex.data_ptr()
        b.footprint_any = self.footprint_any.data_ptr()
        if not mesh_flavour and os.environ.get("LP_BWD_WORKLIST", "1") == "1":
            # this set's forward workspace is its own: the backward may walk the forward's list of live footprints
            wl, wc = ctypes.c_void_p(), ctypes.c_void_p()
            _lib.check(L.lp_forward_worklist(ctypes.byref(a), ctypes.byref(wl), ctypes.byref(wc)))
            b.worklist, b.worklist_ctrl = wl, wc
        if accum is not None:
            # N > 1, fused exchange: the accumulation buffer lives in symmetric memory next to the planar gradient;
            # the backward leaves the gradient interleaved and lp_allreduce_unpack reduces + unpacks + broadcasts it
            self.accum = accum.view(torch.uint8)
            b.workspace, b.workspace_bytes = self.accum.data_ptr(), self.accum.numel()
            b.flags |= _lib.LP_FLAG_GRAD_OVERWRITE | _lib.LP_FLAG_GRAD_INTERLEAVED
        else:
            self.accum = torch.empty(int(L.lp_backward_workspace_bytes(C, T, T)), dtype=torch.uint8, device=device)
            if self.accum.numel() and os.environ.get("LP_BWD_VEC", "1") == "1":
                b.workspace, b.workspace_bytes = self.accum.data_ptr(), self.accum.numel()
                b.flags |= _lib.LP_FLAG_GRAD_OVERWRITE
                if grad_layout == "interleaved":
                    # the gradient stays texel-interleaved (T,T,4) in the accumulation buffer: the layout the fused
                    # optimiser (lp_adam_step with accum) and lp_allreduce_unpack consume; no unpack pass
                    b.flags |= _lib.LP_FLAG_GRAD_INTERLEAVED
            else:
                self.accum = self.accum[:0]
        self.interleaved = bool(b.flags & _lib.LP_FLAG_GRAD_INTERLEAVED)
        if os.environ.get("LP_DEBUG_BWD_STOP"):
            b.flags |= 1 << int(os.environ["LP_DEBUG_BWD_STOP"])
        self.fwd, self.bwd = a, b
        self.launches = 0
        self.launches_prepare = 0
        self.repack()

    def repack(self):
        """(Re)build the texel-interleaved copy of the texture after ``self.tex`` changed."""
        if self.tex_rgba is not None:
            C, T = self.tex.shape[1], self.tex.shape[-1]
            with torch.cuda.device(self.device):
                stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
                _lib.check(_lib.lib().lp_pack_texture(self.tex.data_ptr(), C, T, T, self.tex_rgba.data_ptr(), stream))

    def gradient(self):
        """The texture gradient of the last step as a planar (C,T,T) tensor, whichever layout the step left it in."""
        if self.interleaved:
            C, T = self.tex.shape[1], self.tex.shape[-1]
            return self.accum.view(torch.float32).view(T, T, 4)[:, :, :C].permute(2, 0, 1).contiguous()
        return self.grad_tex

    def prepare(self, stream, with_raster):
        """The texture-independent stages of this set's views on `stream` (a raw cudaStream_t handle):
        geometry + bins, and with `with_raster` also visibility / uv / mask."""
        L = _lib.lib()
        _lib.check(L.lp_render_prepare(ctypes.byref(self.fwd), stream))
        self.launches_prepare = L.lp_last_launch_count()
        if with_raster:
            self.raster(stream)

    def raster(self, stream):
        """Visibility / uv / mask of this set's views (no texture access) on `stream`; needs prepare()."""
        L = _lib.lib()
        _lib.check(L.lp_render_raster(ctypes.byref(self.fwd), stream))
        self.launches_prepare += L.lp_last_launch_count()

    def shade_backward(self, stream, torch_stream, with_raster, aux=None):
        """The rest of the step: (tile rasterizer fused with the texture fetch | texture fetch only), then
        zero the gradient and scatter the upstream gradient into it.  ``aux`` (a torch stream): the accumulation
        buffer is cleared there, NEXT TO the texture fetch instead of between it and the backward (the clear of 16.8 MB is
        otherwise a link of the step's critical chain); the backward is then told not to clear it again."""
        L = _lib.lib()
        aside = aux is not None and self.accum.numel() > 0 and bool(self.bwd.workspace)
        if aside:
            aux.wait_stream(torch_stream)                 # everything that still uses the buffer precedes the clear
            with torch.cuda.stream(aux):
                self.accum.zero_()
        if with_raster:
            _lib.check(L.lp_render_shade(ctypes.byref(self.fwd), stream))
        else:
            _lib.check(render_forward_raster_fused(self.fwd, stream))
        n = L.lp_last_launch_count()
        if not self.accum.numel():
            with torch.cuda.stream(torch_stream):
                self.grad_tex.zero_()
        flags = self.bwd.flags
        if aside:
            torch_stream.wait_stream(aux)
            self.bwd.flags = flags | _lib.LP_FLAG_GRAD_NO_CLEAR
        _lib.check(L.lp_render_backward(ctypes.byref(self.bwd), stream))
        self.bwd.flags = flags
        self.launches = self.launches_prepare + n + L.lp_last_launch_count()

    def run_split(self, stream=None):
        """prepare -> raster -> shade -> backward on one stream: the benched kernels, serially (tests use it)."""
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream) if stream is None else stream
        self.prepare(stream, True)
        self.shade_backward(stream, torch.cuda.current_stream(self.device), True)

    def run(self):
        """Enqueue forward, zero the gradient, backward on torch's current stream."""
        L = _lib.lib()
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(L.lp_render_forward(ctypes.byref(self.fwd), stream))
        n = L.lp_last_launch_count()
        if not self.accum.numel():
            self.grad_tex.zero_()
        _lib.check(L.lp_render_backward(ctypes.byref(self.bwd), stream))
        self.launches = n + L.lp_last_launch_count()


class HostStep:
    """The reference-facing call with HOST buffers: pinned cameras / upstream gradient in, pinned
    image / mask / texture gradient out, through ``lp_render_step_host`` (tests use it too)."""

    def __init__(self, verts, faces, uv, texture, B, H, W, interp, fov, device="cuda:0", flavour="lp"):
        self.device = torch.device(device)
        geom = (verts.to(self.device).float().contiguous(), faces.to(self.device, torch.int32).contiguous(),
                uv.to(self.device).float().reshape(-1, 3, 2).contiguous())
        C, T = texture.shape[1], texture.shape[-1]
        w = dict(B=B, H=H, W=W, C=C, T=T, interp=interp, flavour=flavour)
        self.dev = DeviceStep(geom, w, torch.zeros(B, 4, 3), 0, self.device)
        self.dev.tex.copy_(texture.reshape(1, C, T, T))
        self.dev.repack()
        self.h_cams = torch.empty(B, 4, 3).pin_memory()
        self.h_grad = torch.empty(B, C, H, W).pin_memory()
        self.h_image = torch.empty(B, C, H, W).pin_memory()
        self.h_mask = torch.empty(B, 1, H, W).pin_memory()
        self.h_gtex = torch.empty(C, T, T).pin_memory()
        self.h2d_bytes = self.h_cams.numel() * 4 + self.h_grad.numel() * 4
        self.d2h_bytes = (self.h_image.numel() + self.h_mask.numel() + self.h_gtex.numel()) * 4

    def step_async(self, torch_stream):
        """Enqueue one host-buffer step on `torch_stream` without waiting for it."""
        stream = ctypes.c_void_p(torch_stream.cuda_stream)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().lp_render_step_host_async(
                ctypes.byref(self.dev.fwd), ctypes.byref(self.dev.bwd), self.h_cams.data_ptr(), self.h_grad.data_ptr(),
                self.h_image.data_ptr(), self.h_mask.data_ptr(), self.h_gtex.data_ptr(), stream))

    def step(self, cams=None, grad_image=None):
        if cams is not None:
            self.h_cams.copy_(cams)
        if grad_image is not None:
            self.h_grad.copy_(grad_image)
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().lp_render_step_host(
                ctypes.byref(self.dev.fwd), ctypes.byref(self.dev.bwd), self.h_cams.data_ptr(), self.h_grad.data_ptr(),
                self.h_image.data_ptr(), self.h_mask.data_ptr(), self.h_gtex.data_ptr(), stream))
        return self.h_image, self.h_mask, self.h_gtex


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()

    def run(self):
        # The first query waits 5 ms: spawning nvidia-smi stalls this process (fork of a process with a CUDA context) and
        # the query itself disturbs the GPU for a few hundred microseconds — measured on the 20-step run (1.3 ms timed
        # region): 66 us per step without a query inside it, 81 us with one.  Regions shorter than that are followed by the
        # same steps repeated for half a second (measure()), which is where their clock samples come from.
        self.stop_flag.wait(0.005)
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.samples.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = [float(s[0]) for s in self.samples if len(s) >= 6 and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) >= 6 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples if len(s) >= 6 for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_reference_views_per_s(w, n_views, seed=0, warmup=0, impl="torch"):
    """The oracle port of the reference CPU path: reference glue mirror over the torch kaolin
    restatement (rasterizer in torch, all host threads), forward + backward; one view per call for the
    latent_paint Renderer, batches of up to 8 views for the (batched) latent_paint_mesh Renderer."""
    from oracle import kaolin_shim, renderer_ref
    kaolin_shim.RASTER_IMPL = impl
    torch.set_num_threads(os.cpu_count() or 1)
    verts, faces, uv = load_scene(w)
    tex = (0.4 * torch.randn(1, w["C"], w["T"], w["T"], generator=torch.Generator().manual_seed(seed))).requires_grad_(True)
    flavour = w.get("flavour", "lp")
    radius, theta, phi = make_views(warmup + n_views, seed, flavour)
    t0 = None
    if flavour == "mesh":
        r = renderer_ref.LatentPaintMeshRendererRef(dim=(w["W"], w["H"]))
        per = 8
        i = 0
        while i < warmup + n_views:
            if t0 is None and i >= warmup:
                t0 = time.perf_counter()
            j = min(i + per, warmup + n_views)
            grad = torch.randn(j - i, w["C"], w["H"], w["W"], generator=torch.Generator().manual_seed(seed + 1))
            tex.grad = None
            outs = r.render_single_view_texture(verts, faces, uv, tex, theta[i:j], phi[i:j], radius[i:j],
                                                dims=(w["W"], w["H"]), is_body=True)
            outs[0].backward(grad)
            i = j
    else:
        grad = torch.randn(1, w["C"], w["H"], w["W"], generator=torch.Generator().manual_seed(seed + 1))
        r = renderer_ref.LatentPaintRendererRef(dim=(w["W"], w["H"]), interpolation_mode=w["interp"])
        for i in range(warmup + n_views):
            if i == warmup:
                t0 = time.perf_counter()
            tex.grad = None
            image, _ = r.render_single_view_texture(verts, faces, uv, tex, elev=float(theta[i]), azim=float(phi[i]),
                                                    radius=float(radius[i]), look_at_height=w["dy"])
            image.backward(grad)
    dt = time.perf_counter() - t0
    return n_views / dt, dt


def debug_check(tag):
    """LP_DEBUG_CHECK=1 with a -DLP_CHECKED library: print the device-side violation counter at this point."""
    if os.environ.get("LP_DEBUG_CHECK") != "1":
        return
    torch.cuda.synchronize()
    line = ctypes.c_int32(0)
    print(f"[check] {tag}: {_lib.lib().lp_check_failures(ctypes.byref(line))} violations (first at line {line.value})", file=sys.stderr, flush=True)


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    B = w["B"] // world if w["scaling"] == "strong" else w["B"]
    # each step = one batch of B views on the CPU (≈0.2 s/view on 8 cores for c2)
    total = (args.warmup + args.steps) * B
    cap = 400 if w["H"] <= 512 else 24                  # bound the run to a few minutes
    per_step = B if total <= cap else max(1, cap // (args.warmup + args.steps))
    vps, dt = cpu_reference_views_per_s(w, args.steps * per_step, warmup=args.warmup * per_step)
    line = {"metric": "views/sec (fwd+bwd render)", "value": vps, "unit": "views/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": w["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["label"], "views_per_step": per_step},
            "cpu_baseline": {"value": vps, "unit": "views/s", "cores": os.cpu_count(), "kind": "port",
                             "sample": f"{args.steps} steps x {per_step} views, torch CPU path of the oracle "
                                       f"(reference render.py mirror over the torch kaolin restatement)"},
            "e2e": {"value": vps, "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


class Env:
    """Process-wide state of one bench run: rank / world, device, the process group."""

    def __init__(self):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.device = torch.device("cuda", self.local)
        torch.cuda.set_device(self.device)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.device)
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize(self.device)

    def max_over_ranks(self, x):
        if self.dist is None:
            return float(x)
        t = torch.tensor([float(x)], device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def measure(args, env, w, full):
    """Time ``args.steps`` steps of workload ``w`` (max over ranks).  ``full``: also the roofline, e2e and
    cpu_baseline legs.  Returns the pieces of the JSON line."""
    world, rank, device, dist = env.world, env.rank, env.device, env.dist
    verts, faces, uv = load_scene(w)
    geom = (verts.to(device).float().contiguous(), faces.to(device, torch.int32).contiguous(),
            uv.to(device).float().reshape(-1, 3, 2).contiguous())
    strong = w["scaling"] == "strong"
    if strong and w["B"] % world:
        raise SystemExit(f"bench.py: {w['B']} views do not shard evenly over {world} ranks")
    B = w["B"] // world if strong else w["B"]
    w = dict(w, B=B)                                      # per-rank batch from here on
    H, W, C, T = w["H"], w["W"], w["C"], w["T"]
    V, F = verts.shape[0], faces.shape[0]
    flavour = w.get("flavour", "lp")
    n_sets = args.sets
    sets, symm_bufs, allreduce_mode = [], [], "none" if world == 1 else "nccl"
    allreduce = args.allreduce
    if allreduce == "auto":             # measured: inside the step graph the two-shot peer kernels win at
        allreduce = "multimem" if world >= 4 else "p2p"           # 2 GPUs, the in-switch kernel from 4 GPUs up
    if world > 1 and allreduce != "nccl":
        from latent_nerf_test_b200.parallel import SymmetricGradientBuffer
        want_fused = C <= 4 and os.environ.get("LP_EXCHANGE_FUSED", "1") == "1"
        # attempts, best first: exchange fused with the unpack, then unpack + all-reduce of the planar gradient over
        # symmetric memory, then NCCL.  Each is self-checked against NCCL once and the verdict is agreed on by all ranks.
        for fused in ([True, False] if want_fused else [False]):
            err = None
            try:
                symm_bufs = []
                for s in range(n_sets):
                    sb = SymmetricGradientBuffer(C * T * T, device, interleaved_texels=T * T if fused else 0, channels=C)
                    if allreduce == "p2p":
                        sb.mode = "p2p"
                    elif allreduce == "multimem" and sb.mode != "multimem":
                        raise RuntimeError("no multicast support on this box")
                    symm_bufs.append(sb)
                gen = torch.Generator(device=device).manual_seed(rank)
                if symm_bufs[0].fused:
                    tex4 = torch.randn(T * T, 4, device=device, generator=gen)         # interleaved accumulation buffer
                    symm_bufs[0].accum.view(T * T, 4).copy_(tex4)
                    probe = tex4[:, :C].t().contiguous().reshape(-1)                   # its planar (C, T*T) form
                else:
                    probe = torch.randn(C * T * T, device=device, generator=gen)
                    symm_bufs[0].flat[:probe.numel()].copy_(probe)
                symm_bufs[0].all_reduce()
                dist.all_reduce(probe)
                torch.cuda.synchronize(device)
                if not torch.allclose(symm_bufs[0].flat[:probe.numel()], probe, rtol=1e-5, atol=1e-5):
                    raise RuntimeError("symmetric-memory all-reduce disagrees with NCCL")
            except Exception as exc:
                err = exc
            ok = torch.tensor([0 if err else 1], device=device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()):
                allreduce_mode = symm_bufs[0].mode + (" fused with the unpack" if symm_bufs[0].fused else "")
                break
            print(f"bench.py: {'fused ' if fused else ''}symmetric-memory exchange unavailable ({err}); trying the next form", file=sys.stderr)
            symm_bufs, allreduce_mode = [], "nccl"
    for s in range(n_sets):
        gt = symm_bufs[s].view((C, T, T)) if symm_bufs else None
        acc = symm_bufs[s].accum if symm_bufs and symm_bufs[s].fused else None
        sets.append(DeviceStep(geom, w, workload_cameras(w, B, 1000 * rank + s), 10 * s + 1, device, grad_tex=gt, accum=acc,
                               grad_layout=args.grad_layout))
    set_bytes = sum(t.numel() * t.element_size() for t in (sets[0].tex, sets[0].grad_image, sets[0].image, sets[0].mask,
                                                           sets[0].uv, sets[0].grad_tex))

    # the main stream carries the step's dependent chain (texture fetch -> backward -> exchange): its kernels go first
    # whenever an SM has room (LP_MAIN_PRIORITY=0: default priority)
    stream = torch.cuda.Stream(device, priority=-1 if os.environ.get("LP_MAIN_PRIORITY", "1") == "1" else 0)
    graphs = None
    with torch.cuda.stream(stream):
        for s in sets:                                   # first touch + correctness of the call chain
            s.run()
        stream.synchronize()
        if not args.no_graph:
            try:
                graphs = []
                for s in sets:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=stream):
                        s.run()
                    graphs.append(g)
            except Exception as exc:                      # eager launches still measure the same kernels
                print(f"bench.py: CUDA graph capture failed ({exc}); timing eager launches", file=sys.stderr)
                graphs = None

    # --pipeline: the texture-independent stages of the next step (geometry, bins, visibility, uv) run on a
    # second stream while the current step fetches the texture, back-propagates (and all-reduces); each
    # buffer set has its own workspace and saved-uv buffer
    pipeline = args.pipeline
    if pipeline == "auto":
        pipeline = "deep"
    pipe_deep = pipeline == "deep"
    pipe_raster = pipeline == "raster" or pipe_deep
    pipeline = None if pipeline == "off" else pipeline
    prep_stream = torch.cuda.Stream(device) if pipeline else None
    rast_stream = torch.cuda.Stream(device) if pipe_deep else None
    geom_done = [torch.cuda.Event() for _ in sets]
    prep_done = [torch.cuda.Event() for _ in sets]
    set_free = [torch.cuda.Event() for _ in sets]
    bwd_done = [torch.cuda.Event() for _ in sets]
    pipe_state = {"primed": [False] * len(sets), "prev": None}
    # N > 1: the exchange is bound by NVLink, not by the SMs, so the visibility stage of the NEXT step is released when
    # the backward of this step ends and runs inside the exchange's window — instead of next to the texture fetch and
    # the backward, which are on the step's critical chain (fetch -> backward -> exchange) and slow down when they share
    # the SMs with it (LP_GATE_RASTER=0: free-running)
    gate_raster = world > 1 and os.environ.get("LP_GATE_RASTER", "1") == "1"
    gate_prep = world > 1 and os.environ.get("LP_GATE_PREP", "0") == "1"       # the same for geometry + bins of step k + 2 (measured worse: 134 vs 109 us at 4 GPUs)
    # LP_CLEAR_ASIDE=1: the backward's accumulation buffer is cleared on a fourth stream next to the texture fetch instead
    # of by lp_render_backward itself between the fetch and the scatter (LP_FLAG_GRAD_NO_CLEAR).  Measured slower on
    # config 2 (65.4 / 65.7 vs 63.1 / 63.1 us per step: the clear then competes with the fetch), so it is off by default
    clear_stream = torch.cuda.Stream(device, priority=-1) if pipeline and os.environ.get("LP_CLEAR_ASIDE", "0") == "1" else None
    h_prep = ctypes.c_void_p(prep_stream.cuda_stream) if pipeline else None
    h_rast = ctypes.c_void_p(rast_stream.cuda_stream) if pipe_deep else None
    h_main = ctypes.c_void_p(stream.cuda_stream)

    def exchange(k):
        """The path's one exchange step: sum the texture gradient of set k over the ranks (main stream)."""
        if world > 1:
            if symm_bufs:
                symm_bufs[k].all_reduce()
            else:
                dist.all_reduce(sets[k].grad_tex)

    # N > 1, deep pipeline: the visibility stage of step i + 1 is issued during step i and released by the backward of
    # step i - 1, so it runs in the window of exchange i - 1 (NVLink-bound, SMs idle) and is complete before the texture
    # fetch of step i + 1 asks for it — instead of being released by backward i, which put it in front of that fetch
    # (LP_RASTER_LOOKAHEAD=0: the former schedule; measured at 2 GPUs: 110 us per step, free-running 99 us)
    # (only where the exchange's window is long enough to hold a visibility pass: with the 4 MiB payload of configs[2]
    # the early release only adds contention to the fetch -> backward chain: 57.9 vs 48.5 us per step at 2 GPUs)
    lookahead_default = "1" if 16 * T * T >= (8 << 20) else "0"
    lookahead_on = gate_raster and pipe_deep and os.environ.get("LP_RASTER_LOOKAHEAD", lookahead_default) == "1"
    prep_after_bwd = world > 1 and os.environ.get("LP_PREP_AFTER_BWD", "0") == "1"      # measured worse at 2 GPUs: 94.7 vs 90.5 us per step
    pipe_state["front"] = [False] * len(sets)

    def issue_front(k):
        """Geometry + bins (and visibility / uv) of buffer set k on the side streams."""
        if pipe_state["primed"][k]:
            # the workspace of set k is free again: after its exchange — or, N > 1, already after its backward (the
            # exchange touches the gradient buffers only), which lets this pass start inside the exchange's window
            prep_stream.wait_event(bwd_done[k] if prep_after_bwd else set_free[k])
        if gate_prep and pipe_state["prev"] is not None:
            prep_stream.wait_event(bwd_done[pipe_state["prev"]])
        if pipe_deep:
            sets[k].prepare(h_prep, False)
            geom_done[k].record(prep_stream)
            rast_stream.wait_event(geom_done[k])
            if gate_raster and pipe_state["prev"] is not None:
                rast_stream.wait_event(bwd_done[pipe_state["prev"]])
            sets[k].raster(h_rast)
            prep_done[k].record(rast_stream)
        else:
            sets[k].prepare(h_prep, pipe_raster)
            prep_done[k].record(prep_stream)
        pipe_state["front"][k] = True

    def pipelined_step(i, with_exchange=True, more=False):
        """Step i.  ``more``: step i + 1 follows in the same batch (graph or loop), its front may be issued now."""
        k = i % len(sets)
        if not pipe_state["front"][k]:
            issue_front(k)
        if more and lookahead_on:
            issue_front((i + 1) % len(sets))             # gated on the backward of step i - 1 (pipe_state["prev"])
        pipe_state["front"][k] = False
        stream.wait_event(prep_done[k])
        sets[k].shade_backward(h_main, stream, pipe_raster, aux=clear_stream)
        bwd_done[k].record(stream)
        pipe_state["prev"] = k
        if with_exchange:
            exchange(k)
        set_free[k].record(stream)
        pipe_state["primed"][k] = True
        return k

    # the same pipeline captured once as a CUDA graph of PIPE_STEPS steps (capture streams joined by event edges);
    # one replay = PIPE_STEPS steps; the pipeline drains between replays, so a replay is made long
    PIPE_STEPS = len(sets) * max(2, min(10, args.steps // len(sets)))
    if os.environ.get("LP_PIPE_STEPS"):
        PIPE_STEPS = len(sets) * max(1, int(os.environ["LP_PIPE_STEPS"]) // len(sets))
    pipe_graph, rest_graph = None, None
    REST_STEPS = args.steps % PIPE_STEPS           # a step count that is not a multiple of the replay length: a second,
                                                   # shorter graph for the remainder instead of eager launches
    if pipeline and not args.no_graph and (world == 1 or symm_bufs):
        def capture(n_steps):
            pipe_state["primed"] = [False] * len(sets)
            pipe_state["front"] = [False] * len(sets)
            pipe_state["prev"] = None
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                prep_stream.wait_stream(stream)
                if pipe_deep:
                    rast_stream.wait_stream(stream)
                for i in range(n_steps):
                    pipelined_step(i, more=i + 1 < n_steps)
                stream.wait_stream(prep_stream)
                if pipe_deep:
                    stream.wait_stream(rast_stream)
            pipe_state["primed"] = [False] * len(sets)
            pipe_state["front"] = [False] * len(sets)
            pipe_state["prev"] = None
            return g
        try:
            with torch.cuda.stream(stream):
                debug_check("before the eager pipelined warm-up")
                for i in range(PIPE_STEPS):                     # warm both paths before capture
                    pipelined_step(i)
                stream.wait_stream(prep_stream)
                torch.cuda.synchronize(device)
                debug_check("after the eager pipelined warm-up")
                pipe_graph = capture(PIPE_STEPS)
                if REST_STEPS:
                    rest_graph = capture(REST_STEPS)
        except Exception as exc:
            print(f"bench.py: pipelined graph capture failed ({exc}); running the pipeline eagerly", file=sys.stderr)
            pipe_graph, rest_graph = None, None
            pipe_state["primed"] = [False] * len(sets)
            pipe_state["front"] = [False] * len(sets)
            pipe_state["prev"] = None

    def local_step(i):
        """One step without the exchange (rank-local keep-busy loop)."""
        k = i % len(sets)
        if pipeline:
            return pipelined_step(i, with_exchange=False)
        if graphs is not None:
            graphs[k].replay()
        else:
            sets[k].run()
        return k

    def one_step(i):
        if pipeline:
            pipelined_step(i)
        else:
            exchange(local_step(i))

    def join_side_streams():
        """Order the side streams behind everything on the main stream and forget the pipeline's event state.  A graph
        replay lives on the main stream only: eager pipelined steps that follow one must not start their geometry /
        visibility stages (on the side streams) before the replay's kernels are done with the buffer sets — they did
        when the step count was not a multiple of the replay length (config 4 with --steps 50: two prepare passes of
        one set at once, class counters counted twice, work lists overrun)."""
        if pipeline:
            prep_stream.wait_stream(stream)
            if pipe_deep:
                rast_stream.wait_stream(stream)
            pipe_state["primed"] = [False] * len(sets)
            pipe_state["front"] = [False] * len(sets)
            pipe_state["prev"] = None

    def join_main_stream():
        """The reverse: everything the eager steps left on the side streams precedes what the main stream does next."""
        if pipeline:
            stream.wait_stream(prep_stream)
            if pipe_deep:
                stream.wait_stream(rast_stream)

    def run_steps(n):
        """n steps the way the timed region runs them: whole replays of the pipelined graph, the rest eagerly."""
        if pipe_graph is not None:
            if n >= PIPE_STEPS:
                join_main_stream()
            for _ in range(n // PIPE_STEPS):
                pipe_graph.replay()
            n = n % PIPE_STEPS
            if n and n == REST_STEPS and rest_graph is not None:
                join_main_stream()
                rest_graph.replay()
                n = 0
            if n:
                join_side_streams()
        for i in range(n):
            one_step(i)

    with torch.cuda.stream(stream):
        # warm-up: at least `warmup` steps AND at least one replay of every graph the timed region replays, so the
        # first (slow: upload + first-launch initialisation) replay is outside the timed region
        debug_check("after capture")
        run_steps(max(args.warmup, PIPE_STEPS if pipe_graph is not None else 0))
        if rest_graph is not None:
            join_main_stream()
            rest_graph.replay()
        debug_check("after the warm-up replays")
        if pipeline:
            stream.wait_stream(prep_stream)
        if pipe_deep:
            stream.wait_stream(rast_stream)
        env.barrier()
        sampler = ClockSampler(env.local) if rank == 0 and full else None
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        if pipeline:
            prep_stream.wait_event(e0)
        if pipe_deep:
            rast_stream.wait_event(e0)
        run_steps(args.steps)
        if pipeline:
            stream.wait_stream(prep_stream)
        if pipe_deep:
            stream.wait_stream(rast_stream)
        e1.record(stream)
        env.barrier()
        ms = e0.elapsed_time(e1)
        # keep the GPU busy a little longer if the region was too short for nvidia-smi to sample it
        if sampler and ms < 400:
            t_end = time.time() + 0.5
            join_side_streams()
            while time.time() < t_end:
                for i in range(50):
                    local_step(i)              # no collective here: the other ranks are not in this loop
                torch.cuda.synchronize(device)
        clocks = sampler.summary() if sampler else None
    ms = env.max_over_ranks(ms)
    value = world * B * args.steps / (ms * 1e-3)
    res = {"value": value, "ms": ms, "B": B, "clocks": clocks, "launches_per_step": sets[0].launches,
           "allreduce_mode": allreduce_mode, "pipeline": pipeline or "off",
           "cuda_graph": (graphs is not None and not pipeline) or pipe_graph is not None,
           "l2": f"{len(sets)} rotating buffer sets of {set_bytes / 1e6:.0f} MB each "
                 f"({len(sets) * set_bytes / 1e6:.0f} MB > 126 MB L2): inputs larger than L2"
                 if len(sets) * set_bytes > 126e6 else
                 f"{len(sets)} rotating buffer sets of {set_bytes / 1e6:.1f} MB each ({len(sets) * set_bytes / 1e6:.0f} MB: "
                 f"fits the 126 MB L2 — this workload is L2-resident by nature)",
           "roofline": None, "e2e": None, "cpu": None}
    if not full:
        return res

    # ---- e2e: host buffers through lp_render_step_host, on every rank at once
    if not args.no_e2e:
        # three host-buffer contexts in flight on three streams: the H2D copy of one step, the kernels of the
        # previous and the D2H copy of the one before overlap; every step still moves all its bytes both ways
        n_ctx = 3
        ctxs, streams_e2e = [], [torch.cuda.Stream(device) for _ in range(n_ctx)]
        for j in range(n_ctx):
            hs = HostStep(verts, faces, uv, sets[0].tex.cpu(), B, H, W, w["interp"], FOV, device=str(device), flavour=flavour)
            hs.h_cams.copy_(workload_cameras(w, B, 77 + j + 10 * rank))
            hs.h_grad.copy_(torch.randn(B, C, H, W, generator=torch.Generator().manual_seed(5 + j)))
            ctxs.append(hs)
        n_e2e = max(6, min(args.steps, 60))
        for j in range(2 * n_ctx):
            ctxs[j % n_ctx].step_async(streams_e2e[j % n_ctx])
        env.barrier()
        t0 = time.perf_counter()
        for j in range(n_e2e):
            streams_e2e[j % n_ctx].synchronize()          # the context's previous results have landed on the host
            ctxs[j % n_ctx].step_async(streams_e2e[j % n_ctx])
        torch.cuda.synchronize(device)
        e2e_s = env.max_over_ranks(time.perf_counter() - t0)
        hs = ctxs[0]
        t1 = time.perf_counter()
        for _ in range(5):
            hs.step()
        sync_ms = 1e3 * (time.perf_counter() - t1) / 5
        res["e2e"] = {"value": world * B * n_e2e / e2e_s, "unit": "views/s", "h2d_bytes_per_step": hs.h2d_bytes,
                      "d2h_bytes_per_step": hs.d2h_bytes, "ms_per_step": 1e3 * e2e_s / n_e2e, "n_gpus": world,
                      "bytes_are": "per rank and step", "in_flight": n_ctx, "ms_per_step_one_at_a_time": sync_ms}
        del ctxs
        env.barrier()

    if rank == 0:
        # ---- roofline: instrumented eager pass over the same steps (events around every kernel)
        L = _lib.lib()
        with torch.cuda.stream(stream):
            L.lp_timing_enable(1)
            n_prof = min(args.steps, 200)
            for i in range(n_prof):                 # the kernels of the timed region, one after the other
                if pipe_raster:
                    sets[i % len(sets)].prepare(h_main, True)
                    sets[i % len(sets)].shade_backward(h_main, stream, True)
                else:
                    sets[i % len(sets)].run()
            torch.cuda.synchronize(device)
            timings = _lib.collect_timings()
            L.lp_timing_enable(0)
        per_kernel = {k: 1e3 * v[0] / v[1] for k, v in timings.items()}           # µs per launch
        fwd_bytes, bwd_bytes = algorithmic_bytes(V, F, H, W, C, T, B, flavour)
        if pipe_raster:
            # split forward: the tile kernel writes mask + saved uv only; k_shade reads the uv and the texture and
            # writes the image (its uv read is extra traffic the fused form does not have, so not counted)
            shade_bytes = B * H * W * 4 * C + 4 * C * T * T
            fwd_bytes_tile = fwd_bytes - shade_bytes
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.isfile(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        else:
            peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        kb = {"k_raster_shade": fwd_bytes_tile if pipe_raster else fwd_bytes, "k_backward_texture": bwd_bytes}
        if pipe_raster:
            kb["k_shade"] = shade_bytes
        kb = {k: v for k, v in kb.items() if k in per_kernel}
        dom = max(kb, key=lambda k: per_kernel.get(k, 0.0))
        achieved = kb[dom] / (per_kernel[dom] * 1e-6) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")                       # dram bytes per launch from ncu --set full
        if os.path.isfile(tp) and args.workload == "c2":
            traffic = json.load(open(tp)).get(dom + ("_split" if pipe_raster and dom == "k_raster_shade" else ""))
        res["roofline"] = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                           "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                           "algorithmic_bytes_per_launch": kb[dom], "us_per_launch": per_kernel[dom],
                           "kernels_us": per_kernel,
                           "step": {"algorithmic_bytes": fwd_bytes + bwd_bytes,
                                    "achieved_gbs": (fwd_bytes + bwd_bytes) / (ms * 1e-3 / args.steps) / 1e9,
                                    "frac": (fwd_bytes + bwd_bytes) / (ms * 1e-3 / args.steps) / 1e9 / peak}}

        # ---- cpu baseline: bounded sample of the same workload on the host cores
        if args.cpu_views > 0 and world == 1:
            n_cpu = args.cpu_views if H <= 512 else min(args.cpu_views, 8)
            vps, dt = cpu_reference_views_per_s(w, n_cpu, warmup=2)
            res["cpu"] = {"value": vps, "unit": "views/s", "cores": os.cpu_count(), "kind": "port",
                          "sample": f"{n_cpu} views of the same workload, {dt:.1f} s; oracle torch CPU path"}
    return res


def run_train_step(args):
    """BASELINE.json configs[4]: the slimmed reference training loop of tools/train_step.py (renderer -> guidance UNet ->
    backward -> exchange -> fused Adam), 8 views per GPU."""
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) == 0:
            print(json.dumps({"impl": "reference", "unavailable": "configs[4] times the SD UNet on a GPU; the CPU arm covers the render path (configs[1-3])"}), flush=True)
        return None
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import train_step
    out = train_step.run(views=8, steps=args.steps, warmup=args.warmup, quiet=True)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if out is None:
        return None
    line = {"metric": "views/sec (end-to-end SDS step)", "value": out["views_per_s"], "unit": "views/s", "n_gpus": out["n_gpus"],
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": out["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 renderer / bf16 guidance network", "data": "synthetic",
            "config": {"workload": out["workload"], "views_per_gpu_per_step": 8, "exchange": out["exchange"]},
            "stage_ms": out["ms"], "renderer_ms": out["renderer_ms"], "renderer_share_of_step": out["renderer_share_of_step"],
            "note": out["note"]}
    print(json.dumps(line), flush=True)
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + ["c5"])
    ap.add_argument("--sets", type=int, default=4, help="rotating buffer sets (working set > L2)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--pipeline", default="auto", choices=["auto", "off", "geometry", "raster", "deep"],
                    help="overlap texture-independent stages of later steps with step k on other streams: 'geometry' = setup + "
                         "bins, 'raster' = also visibility/uv (hides the all-reduce when N > 1), 'deep' = three stages on three "
                         "streams (geometry | visibility/uv | texture fetch + backward + exchange); auto = deep")
    ap.add_argument("--allreduce", default="auto", choices=["auto", "nccl", "multimem", "p2p"],
                    help="texture-gradient exchange at N > 1: the library's own NVLink kernels over symmetric memory "
                         "(multimem = NVSwitch in-switch reduction, p2p = two-shot peer loads) or NCCL")
    ap.add_argument("--grad-layout", default="interleaved", choices=["interleaved", "planar"],
                    help="what a step leaves behind: the texture gradient texel-interleaved (T,T,4) — the layout the fused "
                         "optimiser (lp_adam_step) and the exchange (lp_allreduce_unpack) consume — or the reference's planar "
                         "(C,T,T), which costs one more pass (k_unpack_grad)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    ap.add_argument("--no-strong", action="store_true", help="skip the configs[2] strong-scaling measurement of the default run")
    ap.add_argument("--cpu-views", type=int, default=160, help="views timed for cpu_baseline, about 10-30 s of CPU work (0 = skip)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.workload == "c5":
        return run_train_step(args)
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, w)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the render path has no CPU implementation (use --impl reference)")
    env = Env()
    # The persistent footprint kernel needs only two resident CTAs per SM (config 2: 41 us with two, 39 us with four);
    # the pipelined forms leave the rest of each SM to the kernels of the other streams (measured: 73.9 us per step
    # with two, 85.8 with three, 86.8 with four).  LP_RASTER_CTAS / LP_PDL override (experiments).
    if args.pipeline != "off":
        _lib.check(_lib.lib().lp_set_option(_lib.LP_OPT_RASTER_CTAS_PER_SM, 2))
        if env.world == 1:
            # ... and four instead of eight CTAs per SM for the texture fetch and the backward (config 2, four A/B pairs:
            # 61.8-62.5 vs 62.7-63.8 us per step); at N > 1, where that chain is followed by the exchange, the
            # library's default stays
            _lib.check(_lib.lib().lp_set_option(_lib.LP_OPT_WALK_CTAS_PER_SM, 4))
    for name, opt in (("LP_PDL", _lib.LP_OPT_PDL), ("LP_RASTER_CTAS", _lib.LP_OPT_RASTER_CTAS_PER_SM),
                      ("LP_EXCHANGE_CTAS", _lib.LP_OPT_EXCHANGE_CTAS), ("LP_WALK_CTAS", _lib.LP_OPT_WALK_CTAS_PER_SM),
                      ("LP_EXCHANGE_BULK", _lib.LP_OPT_EXCHANGE_BULK)):
        if os.environ.get(name):
            _lib.check(_lib.lib().lp_set_option(opt, int(os.environ[name])))
    res = measure(args, env, w, full=True)
    if os.environ.get("LP_B200_LIB"):                    # a -DLP_CHECKED library counts violated kernel invariants on the device
        line = ctypes.c_int32(0)
        n_bad = _lib.lib().lp_check_failures(ctypes.byref(line))
        if n_bad:
            print(f"bench.py: {n_bad} violated kernel invariants, first at lp_b200.cu:{line.value}", file=sys.stderr)
    strong = None
    if args.workload == "c2" and not args.no_strong and WORKLOADS["c3"]["B"] % env.world == 0:
        # BASELINE.json configs[2] beside the weak-scaling line: 64 teddy views in total, sharded over the ranks
        sres = measure(args, env, WORKLOADS["c3"], full=False)
        strong = {"workload": WORKLOADS["c3"]["label"], "scaling": "strong", "value": sres["value"], "unit": "views/s",
                  "n_gpus": env.world, "views_total": WORKLOADS["c3"]["B"], "views_per_gpu_per_step": sres["B"],
                  "ms_per_step": sres["ms"] / args.steps, "steps": args.steps,
                  "exchange_payload_bytes": 4 * WORKLOADS["c3"]["C"] * WORKLOADS["c3"]["T"] ** 2,
                  "allreduce": sres["allreduce_mode"], "launches_per_step": sres["launches_per_step"]}
    line = None
    if env.rank == 0:
        world = env.world
        line = {"metric": "views/sec (fwd+bwd render)", "value": res["value"], "unit": "views/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms"] / args.steps, "higher_is_better": True,
                "scaling": w["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": w["label"], "views_per_gpu_per_step": res["B"], "cuda_graph": res["cuda_graph"],
                           "pipeline": res["pipeline"], "l2": res["l2"], "gradient_layout": args.grad_layout,
                           "parallelism": f"views sharded over {world} GPU(s)" + (f", all-reduce of the texture gradient each step ({res['allreduce_mode']})" if world > 1 else "")},
                "gpu_launches": res["launches_per_step"] * args.steps, "launches_per_step": res["launches_per_step"],
                "clocks": res["clocks"], "roofline": res["roofline"], "e2e": res["e2e"], "cpu_baseline": res["cpu"],
                "strong_scaling": strong}
        print(json.dumps(line), flush=True)
    if env.dist is not None:
        env.dist.barrier()
        env.dist.destroy_process_group()
    return line


if __name__ == "__main__":
    main()
