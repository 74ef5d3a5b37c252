"""CPU tests of the host side: the C-ABI library builds, loads and exports every symbol the
header declares (no compute calls — there is no GPU here), the ctypes structs match the C
layout, and the host logic (camera math, mesh IO, UV atlas, CSR, view sharding) is right."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import latent_nerf_test_b200 as lp
from latent_nerf_test_b200 import _lib, camera, functional, meshio
from oracle import kaolin_shim as kal
from oracle import renderer_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "lp_b200.h")


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    src = open(HEADER).read()
    declared = sorted(set(re.findall(r"\b(lp_[a-z0-9_]+)\s*\(", src)))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/lp_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == declared
    assert L.lp_version() == 100
    assert L.lp_error_string(3).decode() == "workspace too small"


def test_ctypes_structs_match_c_layout(tmp_path):
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "lp_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n",'
                    'sizeof(LpForwardArgs),sizeof(LpBackwardArgs),offsetof(LpForwardArgs,workspace_bytes),'
                    'offsetof(LpForwardArgs,lights),offsetof(LpBackwardArgs,workspace_bytes));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    assert int(out[0]) == ctypes.sizeof(_lib.LpForwardArgs)
    assert int(out[1]) == ctypes.sizeof(_lib.LpBackwardArgs)
    assert int(out[2]) == _lib.LpForwardArgs.workspace_bytes.offset
    assert int(out[3]) == _lib.LpForwardArgs.lights.offset
    assert int(out[4]) == _lib.LpBackwardArgs.workspace_bytes.offset


def test_python_constants_match_the_header(tmp_path):
    """Flags, options, interpolation modes and the exchange structs of ``_lib.py`` against ``include/lp_b200.h`` (a C
    program prints the header's values)."""
    names = [n for n in dir(_lib) if re.fullmatch(r"LP_(FLAG|OPT|INTERP|ERR)_[A-Z0-9_]+|LP_OK|LP_EXCHANGE_FLAG_BYTES", n)]
    assert {"LP_OPT_WALK_CTAS_PER_SM", "LP_OPT_EXCHANGE_BULK", "LP_FLAG_GRAD_INTERLEAVED", "LP_FLAG_MICRO_ON"} <= set(names)
    body = "".join(f'printf("{n} %lld\\n", (long long)({n}));' for n in names)
    body += 'printf("sizeof_LpExchangeArgs %zu\\n", sizeof(LpExchangeArgs));printf("sizeof_LpAdamArgs %zu\\n", sizeof(LpAdamArgs));'
    prog = tmp_path / "consts.c"
    prog.write_text('#include <stdio.h>\n#include "lp_b200.h"\nint main(){' + body + 'return 0;}\n')
    exe = tmp_path / "consts"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for n in names:
        assert int(out[n]) == int(getattr(_lib, n)), n
    assert int(out["sizeof_LpExchangeArgs"]) == ctypes.sizeof(_lib.LpExchangeArgs)
    assert int(out["sizeof_LpAdamArgs"]) == ctypes.sizeof(_lib.LpAdamArgs)


def test_argument_validation_without_a_gpu():
    """Bad arguments are rejected before any CUDA call is made."""
    L = _lib.lib()
    assert L.lp_render_forward(None, None) == _lib.LP_ERR_BAD_ARG
    a = _lib.LpForwardArgs()
    assert L.lp_render_forward(ctypes.byref(a), None) == _lib.LP_ERR_BAD_ARG
    assert b"verts" in L.lp_last_error()
    assert L.lp_render_backward(None, None) == _lib.LP_ERR_BAD_ARG
    assert L.lp_workspace_bytes(0, 10, 10, 10) == 0
    assert L.lp_workspace_bytes(8, 7500, 512, 512) > 8 * 7500 * 36
    with pytest.raises(ValueError):
        _lib.check(_lib.LP_ERR_BAD_ARG)


def test_renderer_refuses_cpu():
    with pytest.raises(RuntimeError, match="no CPU path"):
        lp.LatentPaintRenderer("cpu")
    with pytest.raises(RuntimeError, match="no CPU path"):
        lp.LatentPaintMeshRenderer("cpu")


def test_host_camera_equals_oracle_camera():
    for e, a, r, h in [(1.0, 0.7, 1.25, 0.25), (0.3, 5.9, 2.0, 0.0), (2.5, 3.1, 1.0, -0.3)]:
        mine = camera.camera_from_view(torch.tensor(e), torch.tensor(a), r, h)
        ref = renderer_ref.LatentPaintRendererRef.get_camera_from_view(torch.tensor(e), torch.tensor(a), r, h)
        assert mine.shape == (1, 4, 3) and torch.equal(mine, ref)
    th, ph, rad = torch.tensor([1.2, 1.7, 1.0]), torch.tensor([0.3, 5.0, 2.2]), torch.tensor([1.5, 2.2, 1.9])
    assert torch.equal(camera.camera_from_view(th, ph, rad, torch.tensor([0.4])),
                       renderer_ref._look_at_camera(th, ph, rad, torch.tensor([0.4])))
    M = camera.camera_from_view(th, ph, rad, 0.4)
    R = M[:, :3, :]
    assert torch.allclose(R.transpose(1, 2) @ R, torch.eye(3).expand(3, 3, 3), atol=1e-6)
    assert torch.equal(camera.generate_perspective_projection(np.pi / 3), kal.generate_perspective_projection(np.pi / 3))


def test_obj_reader_and_packed_meshes(tmp_path):
    p = tmp_path / "t.obj"
    p.write_text("# c\nmtllib x.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nv 0 0 1\nvt 0 0\nvt 1 0\nvt 0 1\nvn 0 0 1\n"
                 "f 1/1/1 2/2/1 3/3/1\nf 1//1 3//1 4//1\nf 2 3 4\n")
    m = meshio.load_obj(str(p))
    assert m.vertices.shape == (4, 3) and m.faces.tolist() == [[0, 1, 2], [0, 2, 3], [1, 2, 3]]
    assert m.face_uvs_idx.tolist() == [[0, 1, 2], [-1, -1, -1], [-1, -1, -1]]
    assert meshio.face_uv_attributes(m).shape == (1, 3, 3, 2)          # falls back to the atlas
    for name, V, F in [("blub", 7106, 14208), ("nascar", 3750, 7500), ("teddy", 2892, 5760), ("sphere", 642, 1280),
                       ("env_sphere", 2562, 5120)]:
        mm = meshio.load_npz(os.path.join(ROOT, "tests", "golden", "meshes", name + ".npz"))
        assert mm.vertices.shape == (V, 3) and mm.faces.shape == (F, 3)
    v = meshio.normalize_vertices(meshio.find_shape("blub").vertices, 0.6, 0.25)
    c = v - torch.tensor([0.0, 0.25, 0.0])
    assert abs(float(c.norm(dim=1).max()) - 0.6) < 1e-6 and float(c.mean(0).abs().max()) < 1e-6


def test_off_reader_mesh_mirror_and_uv_provisioning(tmp_path):
    """The formats either side of the path (SURVEY.md §8 f rank 2): OFF reader, the reference-shaped Mesh class,
    and the UV source selection of init_texture_map with its vt.pth / ft.pth cache format."""
    off = tmp_path / "t.off"
    off.write_text("OFF\n# comment\n4 2 0\n0 0 0\n1 0 0\n0 1 0\n0 0 2\n3 0 1 2\n3 0 2 3\n")
    m = meshio.load_off(str(off))
    assert m.vertices.shape == (4, 3) and m.faces.tolist() == [[0, 1, 2], [0, 2, 3]] and int(m.face_uvs_idx.max()) == -1
    mesh = meshio.Mesh(str(off), "cpu")
    assert mesh.vt.shape[0] == 0 and mesh.ft.min() == -1
    n = mesh.normalize_mesh(target_scale=0.6, dy=0.25)
    assert mesh.vertices is not n.vertices and abs(float((n.vertices - torch.tensor([0.0, 0.25, 0.0])).norm(dim=1).max()) - 0.6) < 1e-6
    st = mesh.standardize_mesh()
    assert abs(float(torch.std(torch.norm(st.vertices, dim=1))) - 1.0) < 1e-5
    with pytest.raises(ValueError, match="not implemented"):
        meshio.Mesh("x.ply", "cpu")

    # 1. complete UVs on the mesh win
    obj = tmp_path / "u.obj"
    obj.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 0 1\nf 1/1 2/2 3/3\n")
    mu = meshio.Mesh(str(obj), "cpu")
    vt, ft = meshio.provision_uvs(mu, cache_dir=tmp_path / "exp")
    assert vt is mu.vt and ft is mu.ft and not (tmp_path / "exp").exists()
    # 3. no UVs, no cache: atlas, written in the reference's cache format (float32 vt, int32 ft, CPU tensors)
    vt, ft = meshio.provision_uvs(mesh, cache_dir=tmp_path / "exp")
    assert vt.shape == (6, 2) and ft.shape == (2, 3) and ft.dtype == torch.int32
    cv, cf = torch.load(tmp_path / "exp" / "vt.pth"), torch.load(tmp_path / "exp" / "ft.pth")
    assert cv.dtype == torch.float32 and cf.dtype == torch.int32 and torch.equal(cv, vt) and torch.equal(cf, ft)
    # 2. the cache wins over a new parametrisation (here: a cache someone else wrote, e.g. the reference's xatlas run)
    torch.save(torch.full((6, 2), 0.5), tmp_path / "exp" / "vt.pth")
    vt2, ft2 = meshio.provision_uvs(mesh, cache_dir=tmp_path / "exp")
    assert float(vt2.min()) == 0.5 and torch.equal(ft2, ft)
    # the kaolin-namespaced readers
    from latent_nerf_test_b200 import kaolin_compat
    kal = kaolin_compat.make_module()
    assert kal.io.off.import_mesh(str(off)).faces.shape == (2, 3) and kal.io.obj.import_mesh(str(obj)).uvs.shape == (3, 2)


def test_grid_atlas_and_subdivision():
    vt, ft = meshio.grid_atlas_uvs(7500)
    assert vt.shape == (22500, 2) and ft.shape == (7500, 3)
    assert float(vt.min()) > 0 and float(vt.max()) < 1
    tri = vt[ft]                                         # (F,3,2); all triangles have the same positive area
    area = 0.5 * ((tri[:, 1, 0] - tri[:, 0, 0]) * (tri[:, 2, 1] - tri[:, 0, 1]) -
                  (tri[:, 2, 0] - tri[:, 0, 0]) * (tri[:, 1, 1] - tri[:, 0, 1]))
    assert torch.allclose(area, area[0].expand_as(area), rtol=1e-3) and float(area[0]) > 0
    s = meshio.subdivide(meshio.find_shape("sphere"), 2)
    assert s.faces.shape == (1280 * 16, 3) and s.vertices.shape == (642 + 1920 + 7680, 3)
    assert torch.allclose(s.vertices.norm(dim=1), torch.ones(s.vertices.shape[0]), atol=1e-5)
    assert meshio.face_uv_attributes(s).shape == (1, 1280 * 16, 3, 2)


def test_vertex_face_csr_order_matches_reference_accumulation():
    faces = meshio.find_shape("sphere").faces.to(torch.int32)
    V = int(faces.max()) + 1
    off, vf = functional.vertex_face_csr(faces, V)
    assert off[0] == 0 and int(off[-1]) == faces.numel()
    fn = torch.randn(2, faces.shape[0], 3, generator=torch.Generator().manual_seed(0))
    ref = renderer_ref.LatentPaintMeshRendererRef.compute_vertex_normals(faces.long(), fn)
    out = torch.zeros(2, V, 3)
    for v in range(V):                                   # the exact loop k_vertex_normals runs
        acc = torch.zeros(2, 3)
        for i in range(int(off[v]), int(off[v + 1])):
            acc = acc + fn[:, int(vf[i])]
        out[:, v] = acc / max(int(off[v + 1]) - int(off[v]), 1)
    assert torch.equal(out, ref)


def test_shard_views():
    from latent_nerf_test_b200.parallel import shard_views
    for n, w in [(64, 8), (8, 8), (5, 4), (3, 8), (0, 2)]:
        spans = [shard_views(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_views(4, 4, 4)


def test_kaolin_compat_surface_and_install():
    """Every kaolin attribute the reference's render path touches exists in the compat tree."""
    import sys
    kc = lp.kaolin_compat.make_module()
    for path in ["render.camera.generate_perspective_projection", "render.camera.generate_transformation_matrix",
                 "render.mesh.prepare_vertices", "render.mesh.rasterize", "render.mesh.dibr_rasterization",
                 "render.mesh.texture_mapping", "render.mesh.spherical_harmonic_lighting",
                 "ops.mesh.index_vertices_by_faces", "ops.mesh.uniform_laplacian", "io.obj.import_mesh"]:
        obj = kc
        for part in path.split("."):
            obj = getattr(obj, part)
        assert callable(obj), path
    # small ops agree with the oracle on CPU (the heavy ones need the GPU)
    v = torch.randn(5, 3, generator=torch.Generator().manual_seed(0))
    f = torch.tensor([[0, 1, 2], [2, 3, 4]])
    M = camera.camera_from_view(torch.tensor(1.0), torch.tensor(0.5), 1.5, 0.1)
    proj = camera.generate_perspective_projection(np.pi / 3)
    a = kc.render.mesh.prepare_vertices(v, f, proj, camera_transform=M)
    b = kal.prepare_vertices(v, f, proj, camera_transform=M)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    n = torch.randn(1, 4, 4, 3, generator=torch.Generator().manual_seed(1))
    L = torch.tensor([[1.0, 0.0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0]])
    assert torch.equal(kc.render.mesh.spherical_harmonic_lighting(n, L), kal.spherical_harmonic_lighting(n, L))
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == "kaolin" or k.startswith("kaolin.")}
    try:
        for k in saved:
            del sys.modules[k]
        lp.kaolin_compat.install()
        import kaolin as kk
        assert kk.render.mesh.rasterize is lp.kaolin_compat.rasterize
    finally:
        for k in [k for k in sys.modules if k == "kaolin" or k.startswith("kaolin.")]:
            del sys.modules[k]
        sys.modules.update({k: v for k, v in saved.items() if v is not None})


def test_micro_face_sign_pretest_never_rejects_a_covered_pixel():
    """k_setup_count skips the decree's divisions when `w_k * sign(s) < -|s| * 1e-30` for some k (csrc/lp_b200.cu,
    micro-face path).  That shortcut must never reject a pixel the exact test `w_k / s >= 0` accepts — including
    quotients that underflow to -0 (which the decree accepts), zeros, infinities and NaNs.  numpy float32 follows the
    same IEEE rules as the kernel compiled with -fmad=false."""
    rng = np.random.default_rng(0)
    n = 2_000_000
    mag = np.float32(10.0) ** rng.uniform(-44, 38, n).astype(np.float32)
    w = (mag * rng.choice(np.array([-1.0, 1.0], np.float32), n)).astype(np.float32)
    s = (np.float32(10.0) ** rng.uniform(-8, 38, n).astype(np.float32) * rng.choice(np.array([-1.0, 1.0], np.float32), n)).astype(np.float32)
    special = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-45, -1e-45, 1e-38, -1e-38, 3e38, -3e38, 1e-8, -1e-8], np.float32)
    w = np.concatenate([w, np.repeat(special, len(special))])
    s = np.concatenate([s, np.tile(special, len(special))])
    with np.errstate(all="ignore"):
        sg = np.copysign(np.float32(1.0), s)
        guard = (np.abs(s) * np.float32(1e-30)).astype(np.float32)
        rejected = (w * sg).astype(np.float32) < -guard
        accepted_exact = (w / s).astype(np.float32) >= 0
    assert not np.any(rejected & accepted_exact)
    assert rejected.mean() > 0.3                     # and the shortcut does fire on ordinary negative quotients


def test_depth_face_key_order():
    """The 64-bit key of the micro-face path, `orderable(z0) << 32 | (0xFFFFFFFF - face)`: its maximum must be
    "largest z0, ties to the lowest face id", and the float <-> uint map must round-trip."""
    z = np.array([-np.inf, -3e38, -2.5, -1.0, -1e-30, -1e-45, 0.0, 1e-45, 1e-30, 1.0, 7.5, 3e38, np.inf], np.float32)
    u = z.view(np.uint32)
    o = np.where(u & 0x80000000, ~u, u | 0x80000000).astype(np.uint32)
    assert np.all(np.diff(o.astype(np.int64)) > 0)                     # strictly increasing with z
    back = np.where(o & 0x80000000, o ^ 0x80000000, ~o).astype(np.uint32).view(np.float32)
    assert np.array_equal(back, z)
    assert o.min() > 0                                                  # key 0 is free to mean "no face"
    rng = np.random.default_rng(1)
    for _ in range(200):
        zs = rng.choice(z[1:-1], 6).astype(np.float32)
        fs = rng.permutation(1000)[:6].astype(np.uint32)
        us = zs.view(np.uint32)
        os_ = np.where(us & 0x80000000, ~us, us | 0x80000000).astype(np.uint64)
        keys = (os_ << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - fs.astype(np.uint64))
        k = int(np.argmax(keys))
        best = max(zs)
        assert zs[k] == best and fs[k] == min(f for zz, f in zip(zs, fs) if zz == best)


def test_host_side_errors_of_the_widened_rows():
    """The rows added per SURVEY.md §8(f) fail loudly on the host before touching a device."""
    with pytest.raises(RuntimeError, match="no CPU path"):
        lp.optim.FusedAdam([torch.zeros(8)])
    with pytest.raises(ValueError, match="linear_rgb_estimator"):
        lp.textured_mesh.render_train(None, None, None, None, None, torch.zeros(1, 4, 3, 4), 1.0, 0.5, 1.25, latent_mode=False)
    with pytest.raises(RuntimeError, match="no CPU path"):
        lp.LatentPaintRenderer("cpu")


def test_workspace_size_is_a_pure_host_function():
    """lp_workspace_bytes: per-(view, face) records + bins + the 8-byte-per-pixel key buffer of the micro-face path
    (always reserved: whether a call takes that path depends on its flags); monotone in every argument; bad sizes give 0."""
    L = _lib.lib()
    B, H, W = 2, 64, 64
    a, b = L.lp_workspace_bytes(B, 255, H, W), L.lp_workspace_bytes(B, 256, H, W)
    per_face = 3 * 16 + 2 * 16 + 4                                                                # vertex records, edge tests
    assert b >= a >= B * 255 * per_face + 8 * B * H * W
    assert L.lp_workspace_bytes(2 * B, 255, H, W) > a and L.lp_workspace_bytes(B, 255, 2 * H, W) > a
    assert L.lp_workspace_bytes(0, 10, 8, 8) == 0 and L.lp_workspace_bytes(1, 10, 0, 8) == 0
    assert L.lp_backward_workspace_bytes(3, 1024, 1024) == 16 * 1024 * 1024 and L.lp_backward_workspace_bytes(5, 8, 8) == 0


def test_algorithmic_bytes_match_the_survey():
    """bench.py's roofline numerator is SURVEY.md §8(d)'s formula: config 2 = 119.96 MB per 8-view step."""
    import bench
    fwd, bwd = bench.algorithmic_bytes(3750, 7500, 512, 512, 3, 1024, 8)
    assert fwd + bwd == 8 * (12 * 3750 + 36 * 7500 + 512 * 512 * (8 * 3 + 20)) + 8 * 3 * 1024 * 1024 == 119960512
    assert abs((fwd + bwd) / 6461.5e9 * 1e6 - 18.57) < 0.01                                      # µs at the measured HBM peak


def test_sparse_uniform_laplacian_equals_the_dense_one():
    """SURVEY.md 8 f rank 4: the CSR uniform Laplacian replaces the reference's dense V x V matrix
    (latent_paint_mesh textured_mesh.py:60-71); same operator, same Laplacian coordinates, and it scales to config 4."""
    from latent_nerf_test_b200 import kaolin_compat as kc
    for name in ("sphere", "teddy"):
        m = meshio.find_shape(name)
        V = m.vertices.shape[0]
        dense = kc.uniform_laplacian(V, m.faces)
        sparse = kc.uniform_laplacian_sparse(V, m.faces)
        assert sparse.layout == torch.sparse_csr and sparse.shape == (V, V)
        assert torch.allclose(sparse.to_dense(), dense, atol=1e-7)
        assert torch.equal(kal.uniform_laplacian(V, m.faces), dense)                     # the oracle's dense form
        lc_d, lc_s = kc.laplacian_coordinates(dense, m.vertices), kc.laplacian_coordinates(sparse, m.vertices)
        assert torch.allclose(lc_d, lc_s, atol=1e-5)
        disp = 0.01 * torch.randn(V, 3, generator=torch.Generator().manual_seed(0))
        assert torch.allclose(kc.lap_loss(dense, m.vertices + disp, lc_d), kc.lap_loss(sparse, m.vertices + disp, lc_s), rtol=1e-4)
    big = meshio.subdivide(meshio.find_shape("sphere"), 4)                               # 163 842 vertices: dense would be 107 GB
    Ls = kc.uniform_laplacian_sparse(big.vertices.shape[0], big.faces)
    assert Ls.values().numel() <= 8 * big.vertices.shape[0]
    assert float(kc.laplacian_coordinates(Ls, big.vertices).norm(dim=1).max()) < 0.01   # a smooth sphere
