"""N > 1 path on CPU: view sharding + one all-reduce of the flat gradient bucket (gloo, world_size 2).
The render itself is the CPU oracle here (the CUDA path is covered by the -m gpu tests); what is under test
is the host logic of latent-nerf-test_b200/parallel.py: every view rendered exactly once, the summed
gradient equal to the single-process gradient over the whole batch."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.common import assert_close, mesh_views, rnd, scene


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _full_batch_grad(B):
    from oracle import renderer_ref
    verts, faces, uv = scene("sphere", 1.0, 0.0)
    radius, theta, phi = mesh_views(B, seed=4)
    tex = rnd((1, 4, 32, 32), 1).requires_grad_(True)
    g = rnd((B, 4, 24, 24), 2)
    r = renderer_ref.LatentPaintMeshRendererRef(dim=(24, 24))
    image, *_ = r.render_single_view_texture(verts, faces, uv, tex, theta, phi, radius, dims=(24, 24))
    image.backward(g)
    return tex.grad.clone()


def _worker(rank, world, port, B, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from latent_nerf_test_b200.parallel import GradientBucket, render_views_sharded
        from oracle import renderer_ref
        verts, faces, uv = scene("sphere", 1.0, 0.0)
        radius, theta, phi = mesh_views(B, seed=4)
        tex = torch.nn.Parameter(rnd((1, 4, 32, 32), 1))
        extra = torch.nn.Parameter(torch.zeros(7))                    # a second learnable (e.g. background colours)
        bucket = GradientBucket([tex, extra])
        g = rnd((B, 4, 24, 24), 2)
        r = renderer_ref.LatentPaintMeshRendererRef(dim=(24, 24))

        def render_fn(elev, azim, radius, grad):
            image, *_ = r.render_single_view_texture(verts, faces, uv, tex, elev, azim, radius, dims=(24, 24))
            image.backward(grad)
            return image.detach()

        bucket.zero_()
        out, (lo, hi) = render_views_sharded(render_fn, dict(elev=theta, azim=phi, radius=radius, grad=g), B)
        assert tex.grad.data_ptr() == bucket.flat.data_ptr()          # gradients accumulate inside the bucket
        covered = torch.zeros(B)
        covered[lo:hi] = 1
        dist.all_reduce(covered)
        assert torch.equal(covered, torch.ones(B)), "every view must be rendered exactly once"
        assert (out is None) == (hi == lo)
        bucket.all_reduce()
        if rank == 0:
            torch.save(tex.grad.clone(), out_path)
    finally:
        dist.destroy_process_group()


def test_sharded_views_allreduce_matches_single_process(tmp_path):
    B, world = 5, 2                                                   # uneven split: 3 + 2 views
    out_path = str(tmp_path / "grad.pt")
    mp.spawn(_worker, args=(world, _free_port(), B, out_path), nprocs=world, join=True)
    assert_close(torch.load(out_path), _full_batch_grad(B), "all-reduced texture gradient", rtol=1e-4, atol=1e-5)


def test_more_ranks_than_views(tmp_path):
    out_path = str(tmp_path / "grad.pt")
    mp.spawn(_worker, args=(2, _free_port(), 1, out_path), nprocs=2, join=True)
    assert_close(torch.load(out_path), _full_batch_grad(1), "gradient with an idle rank", rtol=1e-4, atol=1e-5)
