"""ctypes binding of ``liblp_b200.so`` (the C ABI declared in ``include/lp_b200.h``).

The library is built in-tree with nvcc for sm_100a (``build()``).  There is no fallback: if
the shared object is missing and cannot be built, importing the binding raises.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
from ctypes import POINTER, Structure, c_char_p, c_float, c_int32, c_uint32, c_uint64, c_void_p

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
SRC = os.path.join(PKG_DIR, "csrc", "lp_b200.cu")
HEADER = os.path.join(ROOT, "include", "lp_b200.h")
LIB_PATH = os.path.join(PKG_DIR, "liblp_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-fmad=false",  # the visibility path must not contract a*b+c (see csrc/lp_b200.cu header)
              "-Xcompiler", "-fPIC", "-shared"]

LP_OK, LP_ERR_BAD_ARG, LP_ERR_UNSUPPORTED, LP_ERR_WORKSPACE, LP_ERR_CUDA = 0, 1, 2, 3, 4
LP_INTERP_NEAREST, LP_INTERP_BILINEAR, LP_INTERP_BICUBIC = 0, 1, 2
LP_FLAG_MASK_IMAGE = 1 << 0
LP_FLAG_WHITE_BACKGROUND = 1 << 1
LP_FLAG_REJECT_BEHIND = 1 << 2
LP_FLAG_CULL_NZ_ZERO = 1 << 3
LP_FLAG_SHADE_FEATURES = 1 << 4
LP_FLAG_GRAD_OVERWRITE = 1 << 5
LP_FLAG_GRAD_INTERLEAVED = 1 << 6
LP_FLAG_BBOX_HALF_OPEN = 1 << 7
LP_FLAG_PLAIN_EPS = 1 << 8
LP_FLAG_AFFINE_INTERP = 1 << 9
LP_FLAG_SH_BAND1_XZY = 1 << 10
LP_FLAG_GRAD_NO_CLEAR = 1 << 11
LP_FLAG_MICRO_OFF = 1 << 22
LP_FLAG_MICRO_ON = 1 << 23
LP_OPT_PDL = 1
LP_OPT_RASTER_CTAS_PER_SM = 2
LP_OPT_EXCHANGE_CTAS = 3
LP_OPT_WALK_CTAS_PER_SM = 4
LP_OPT_EXCHANGE_BULK = 5

EXPORTS = ["lp_version", "lp_last_error", "lp_error_string", "lp_workspace_bytes", "lp_cameras_from_views",
           "lp_render_forward", "lp_render_backward", "lp_vertex_normals", "lp_render_step_host",
           "lp_last_launch_count", "lp_timing_enable", "lp_timing_collect", "lp_backward_workspace_bytes", "lp_texture_map_forward", "lp_render_prepare", "lp_render_raster", "lp_render_shade", "lp_render_raster_shade", "lp_allreduce_multimem", "lp_allreduce_p2p", "lp_allreduce_unpack", "lp_adam_step", "lp_render_step_host_async", "lp_set_option", "lp_pack_texture", "lp_exchange_step", "lp_resize_bicubic", "lp_check_failures", "lp_forward_worklist", "lp_debug_trace"]


class LpForwardArgs(Structure):
    _fields_ = [
        ("verts", c_void_p), ("faces", c_void_p), ("V", c_int32), ("F", c_int32),
        ("face_vertices_image", c_void_p), ("face_vertices_z", c_void_p), ("valid_faces", c_void_p),
        ("cameras", c_void_p), ("B", c_int32), ("proj", c_float * 3), ("H", c_int32), ("W", c_int32),
        ("multiplier", c_float), ("eps", c_float), ("flags", c_uint32),
        ("face_uv", c_void_p), ("texture", c_void_p), ("C", c_int32), ("Th", c_int32), ("Tw", c_int32),
        ("interp", c_int32),
        ("face_features", c_void_p), ("D", c_int32), ("features_batched", c_int32),
        ("vf_offsets", c_void_p), ("vf_faces", c_void_p), ("face_normals", c_void_p), ("vertex_normals", c_void_p),
        ("lights", c_void_p),
        ("image", c_void_p), ("mask", c_void_p), ("uv", c_void_p), ("face_idx", c_void_p), ("bary", c_void_p),
        ("depth", c_void_p), ("normals", c_void_p), ("lighting", c_void_p), ("footprint_any", c_void_p),
        ("workspace", c_void_p), ("workspace_bytes", c_uint64),
        ("under_image", c_void_p), ("under_mask", c_void_p), ("composed", c_void_p),
        ("texture_rgba", c_void_p),
    ]


class LpBackwardArgs(Structure):
    _fields_ = [
        ("B", c_int32), ("H", c_int32), ("W", c_int32), ("flags", c_uint32),
        ("grad_image", c_void_p), ("uv", c_void_p),
        ("C", c_int32), ("Th", c_int32), ("Tw", c_int32), ("interp", c_int32),
        ("grad_texture", c_void_p), ("grad_texture_batch_stride", ctypes.c_int64),
        ("face_idx", c_void_p), ("bary", c_void_p),
        ("F", c_int32), ("D", c_int32), ("features_batched", c_int32),
        ("grad_face_features", c_void_p), ("footprint_any", c_void_p),
        ("workspace", c_void_p), ("workspace_bytes", c_uint64),
        ("under_mask", c_void_p),
        ("worklist", c_void_p), ("worklist_ctrl", c_void_p),
    ]


class LpAdamArgs(Structure):
    _fields_ = [
        ("accum", c_void_p), ("grad", c_void_p), ("param", c_void_p), ("exp_avg", c_void_p), ("exp_avg_sq", c_void_p),
        ("ntex", ctypes.c_int64), ("C", c_int32), ("lr", c_float), ("beta1", c_float), ("beta2", c_float), ("eps", c_float),
        ("step", c_int32),
    ]


class LpExchangeArgs(Structure):
    _fields_ = [
        ("multicast_base", c_void_p), ("buffer_ptrs_dev", c_void_p),
        ("accum_offset", c_uint64), ("grad_offset", c_uint64), ("flags_offset", c_uint64),
        ("ntex", ctypes.c_int64), ("C", c_int32), ("rank", c_int32), ("world", c_int32), ("adam", c_int32),
        ("param_offset", c_uint64), ("exp_avg", c_void_p), ("exp_avg_sq", c_void_p),
        ("lr", c_float), ("beta1", c_float), ("beta2", c_float), ("eps", c_float), ("step", c_int32),
    ]


LP_EXCHANGE_FLAG_BYTES = 8192


class LpResizeArgs(Structure):
    _fields_ = [("inp", c_void_p * 8), ("out", c_void_p * 8), ("planes", c_int32 * 8),
                ("n", c_int32), ("H", c_int32), ("W", c_int32), ("OH", c_int32), ("OW", c_int32), ("backward", c_int32)]


class LpTextureMapArgs(Structure):
    _fields_ = [
        ("B", c_int32), ("H", c_int32), ("W", c_int32), ("uv", c_void_p), ("texture", c_void_p),
        ("texture_batch_stride", ctypes.c_int64), ("C", c_int32), ("Th", c_int32), ("Tw", c_int32), ("interp", c_int32),
        ("out", c_void_p),
    ]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile ``csrc/lp_b200.cu`` into ``liblp_b200.so`` next to this file (in-tree, so the
    object travels to the GPU box).  nvcc cross-compiles sm_100a without a GPU."""
    newest_src = max(os.path.getmtime(SRC), os.path.getmtime(HEADER))
    if not force and os.path.isfile(LIB_PATH) and os.path.getmtime(LIB_PATH) >= newest_src:
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(nvcc):
        raise RuntimeError("liblp_b200.so is not built and nvcc was not found; run __graft_entry__.build() where CUDA is installed")
    cmd = [nvcc, *NVCC_FLAGS, "-I", os.path.join(ROOT, "include"), "-o", LIB_PATH, SRC]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    """Load (building first if the sources are newer) and type the C ABI."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("LP_B200_LIB") or build()      # LP_B200_LIB: load a specific build (kernel A/B experiments)
    L = ctypes.CDLL(path)
    L.lp_version.restype = c_int32
    L.lp_last_error.restype = c_char_p
    L.lp_error_string.restype = c_char_p
    L.lp_error_string.argtypes = [c_int32]
    L.lp_workspace_bytes.restype = c_uint64
    L.lp_workspace_bytes.argtypes = [c_int32, c_int32, c_int32, c_int32]
    L.lp_backward_workspace_bytes.restype = c_uint64
    L.lp_backward_workspace_bytes.argtypes = [c_int32, c_int32, c_int32]
    L.lp_cameras_from_views.restype = c_int32
    L.lp_cameras_from_views.argtypes = [c_void_p, c_void_p, c_void_p, c_int32, c_float, c_int32, c_void_p, c_void_p]
    L.lp_render_forward.restype = c_int32
    L.lp_render_forward.argtypes = [POINTER(LpForwardArgs), c_void_p]
    for name in ("lp_render_prepare", "lp_render_raster", "lp_render_shade", "lp_render_raster_shade"):
        getattr(L, name).restype = c_int32
        getattr(L, name).argtypes = [POINTER(LpForwardArgs), c_void_p]
    L.lp_allreduce_multimem.restype = c_int32
    L.lp_allreduce_multimem.argtypes = [c_void_p, ctypes.c_int64, c_int32, c_int32, c_void_p]
    L.lp_allreduce_p2p.restype = c_int32
    L.lp_allreduce_p2p.argtypes = [c_void_p, ctypes.c_int64, c_int32, c_int32, c_int32, c_void_p]
    L.lp_adam_step.restype = c_int32
    L.lp_adam_step.argtypes = [POINTER(LpAdamArgs), c_void_p]
    L.lp_allreduce_unpack.restype = c_int32
    L.lp_allreduce_unpack.argtypes = [c_void_p, c_void_p, c_uint64, c_uint64, ctypes.c_int64, c_int32, c_int32, c_int32, c_void_p]
    L.lp_render_backward.restype = c_int32
    L.lp_render_backward.argtypes = [POINTER(LpBackwardArgs), c_void_p]
    L.lp_texture_map_forward.restype = c_int32
    L.lp_texture_map_forward.argtypes = [POINTER(LpTextureMapArgs), c_void_p]
    L.lp_vertex_normals.restype = c_int32
    L.lp_vertex_normals.argtypes = [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p]
    L.lp_render_step_host.restype = c_int32
    L.lp_render_step_host.argtypes = [POINTER(LpForwardArgs), POINTER(LpBackwardArgs), c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p]
    L.lp_render_step_host_async.restype = c_int32
    L.lp_render_step_host_async.argtypes = L.lp_render_step_host.argtypes
    L.lp_last_launch_count.restype = c_int32
    L.lp_timing_enable.restype = c_int32
    L.lp_timing_enable.argtypes = [c_int32]
    L.lp_pack_texture.restype = c_int32
    L.lp_pack_texture.argtypes = [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p]
    L.lp_exchange_step.restype = c_int32
    L.lp_exchange_step.argtypes = [POINTER(LpExchangeArgs), c_void_p]
    L.lp_resize_bicubic.restype = c_int32
    L.lp_resize_bicubic.argtypes = [POINTER(LpResizeArgs), c_void_p]
    L.lp_check_failures.restype = c_int32
    L.lp_check_failures.argtypes = [POINTER(c_int32)]
    L.lp_forward_worklist.restype = c_int32
    L.lp_forward_worklist.argtypes = [POINTER(LpForwardArgs), POINTER(c_void_p), POINTER(c_void_p)]
    L.lp_set_option.restype = c_int32
    L.lp_set_option.argtypes = [c_int32, c_int32]
    L.lp_timing_collect.restype = c_int32
    L.lp_timing_collect.argtypes = [c_int32, POINTER(c_char_p), POINTER(c_float), POINTER(c_int32)]
    _lib = L
    return L


def collect_timings() -> dict:
    """{kernel name: (total_ms, launches)} recorded since ``lp_timing_enable(1)``."""
    n = 32
    names, ms, cnt = (c_char_p * n)(), (c_float * n)(), (c_int32 * n)()
    k = lib().lp_timing_collect(n, names, ms, cnt)
    return {names[i].decode(): (float(ms[i]), int(cnt[i])) for i in range(k)}


def check(rc: int) -> None:
    """Translate an LP_ERR_* code into the Python exception the reference's callers would see
    from kaolin/torch (ValueError for bad arguments, RuntimeError otherwise)."""
    if rc == LP_OK:
        return
    msg = lib().lp_last_error().decode()
    if rc in (LP_ERR_BAD_ARG, LP_ERR_UNSUPPORTED):
        raise ValueError(f"lp_b200: {msg}")
    raise RuntimeError(f"lp_b200 ({lib().lp_error_string(rc).decode()}): {msg}")
