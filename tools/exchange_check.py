"""N-GPU check of the exchange kernels (run under torchrun): the one-launch exchange (handshakes inside the kernel) against
NCCL, repeated back to back and replayed from a CUDA graph, its sharded-Adam epilogue against torch.optim.Adam on the summed
gradient, and the time per exchange of every form.  Prints one JSON object on rank 0."""
import json, os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import latent_nerf_test_b200 as lp
from latent_nerf_test_b200.parallel import SymmetricGradientBuffer

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
res = {"world": world}
for C, T in ((3, 1024), (4, 512)):
    ntex = T * T
    sb = SymmetricGradientBuffer(C * ntex, dev, interleaved_texels=ntex, channels=C, with_params=True)
    if os.environ.get("LP_CHECK_MODE"):                  # p2p: the peer form (bulk asynchronous copies) where multicast is the default
        sb.mode = os.environ["LP_CHECK_MODE"]
    res.setdefault("mode", sb.mode)
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    ok = True
    for it in range(5):                                  # back to back: epochs, buffer reuse
        acc = torch.randn(ntex, 4, device=dev, generator=gen)
        sb.accum.view(ntex, 4).copy_(acc)
        ref = acc[:, :C].t().contiguous()
        sb.exchange_step()
        dist.all_reduce(ref)
        torch.cuda.synchronize()
        ok = ok and torch.allclose(sb.flat[:C * ntex].view(C, ntex), ref, rtol=1e-5, atol=1e-5)
    res[f"one_launch_equals_nccl_C{C}_T{T}"] = bool(ok)
    # replayed from a CUDA graph (the epoch must advance inside the kernel, not through its arguments)
    s = torch.cuda.Stream(dev)
    with torch.cuda.stream(s):
        sb.exchange_step(); s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            sb.exchange_step()
        okg = True
        for it in range(4):
            acc = torch.randn(ntex, 4, device=dev, generator=gen)
            sb.accum.view(ntex, 4).copy_(acc)
            ref = acc[:, :C].t().contiguous()
            g.replay()
            dist.all_reduce(ref)
            s.synchronize(); torch.cuda.synchronize()
            okg = okg and torch.allclose(sb.flat[:C * ntex].view(C, ntex), ref, rtol=1e-5, atol=1e-5)
    res[f"graph_replay_equals_nccl_C{C}_T{T}"] = bool(okg)
    # sharded Adam in the epilogue vs torch.optim.Adam on the NCCL-summed gradient
    p0 = 0.4 * torch.randn(C, ntex, device=dev, generator=torch.Generator(device=dev).manual_seed(7))
    sb.params[:C * ntex].view(C, ntex).copy_(p0)
    pref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([pref], lr=0.01, betas=(0.9, 0.99), eps=1e-15)
    dist.barrier(); torch.cuda.synchronize()
    oka = True
    for it in range(3):
        acc = torch.randn(ntex, 4, device=dev, generator=gen) * (10.0 ** (it - 1))
        sb.accum.view(ntex, 4).copy_(acc)
        gsum = acc[:, :C].t().contiguous(); dist.all_reduce(gsum)
        sb.exchange_step(adam=dict(lr=0.01, betas=(0.9, 0.99), eps=1e-15))
        pref.grad = gsum; opt.step()
        torch.cuda.synchronize()
        oka = oka and torch.allclose(sb.params[:C * ntex].view(C, ntex), pref.detach(), rtol=1e-5, atol=1e-6)
    res[f"sharded_adam_equals_torch_C{C}_T{T}"] = bool(oka)
    # timing (events on the stream, 50 exchanges each)
    def timed(fn, n=50):
        for _ in range(5): fn()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n * 1e3], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    res[f"us_one_launch_C{C}_T{T}"] = timed(sb.exchange_step)
    for n in (74, 296):
        lp._lib.lib().lp_set_option(lp._lib.LP_OPT_EXCHANGE_CTAS, n)
        res[f"us_one_launch_{n}ctas_C{C}_T{T}"] = timed(sb.exchange_step)
    lp._lib.lib().lp_set_option(lp._lib.LP_OPT_EXCHANGE_CTAS, 0)
    res[f"us_one_launch_adam_C{C}_T{T}"] = timed(lambda: sb.exchange_step(adam=dict(lr=0.01, betas=(0.9, 0.99), eps=1e-15)))
    if sb.mode == "multimem":                            # the same kernel over plain peer pointers
        sb.mode = "p2p"
        res[f"us_one_launch_p2p_C{C}_T{T}"] = timed(sb.exchange_step)
        sb.mode = "multimem"
    sb.one_launch = False
    res[f"us_barrier_kernel_barrier_C{C}_T{T}"] = timed(sb.all_reduce)
    flat = torch.randn(C * ntex, device=dev)
    res[f"us_nccl_C{C}_T{T}"] = timed(lambda: dist.all_reduce(flat))
    del sb
if rank == 0:
    print(json.dumps(res))
dist.barrier(); dist.destroy_process_group()
