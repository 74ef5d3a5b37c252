"""Device time of the texture-gradient all-reduce variants (run under torchrun on N GPUs of one box)."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from latent_nerf_test_b200.parallel import SymmetricGradientBuffer
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3 * 1024 * 1024
from latent_nerf_test_b200 import _lib
if os.environ.get("LP_EXCHANGE_BULK"):         # 0: register loads instead of bulk asynchronous copies
    _lib.check(_lib.lib().lp_set_option(_lib.LP_OPT_EXCHANGE_BULK, int(os.environ["LP_EXCHANGE_BULK"])))
if os.environ.get("LP_EXCHANGE_CTAS"):         # CTAs of the one-launch exchange (default: one per SM)
    _lib.check(_lib.lib().lp_set_option(_lib.LP_OPT_EXCHANGE_CTAS, int(os.environ["LP_EXCHANGE_CTAS"])))
sb = SymmetricGradientBuffer(n, dev)
plain = torch.zeros(n, device=dev)
def timed(fn, iters=200):
    for _ in range(20): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
res = {"nccl": timed(lambda: dist.all_reduce(plain))}
modes = ["multimem", "p2p"] if sb.multicast_ptr else ["p2p"]
for m in modes:
    sb.mode = m
    res[m] = timed(sb.all_reduce)
# the exchange fused with the unpack (3 channels over 1024^2 texels: 16.8 MB interleaved in, 12.6 MB planar out)
try:
    T = 1024
    fb = SymmetricGradientBuffer(3 * T * T, dev, interleaved_texels=T * T, channels=3)
    for m in (["multimem", "p2p"] if fb.multicast_ptr else ["p2p"]):
        fb.mode = m
        res["fused_unpack_" + m] = timed(fb.all_reduce)
    fb.mode = "multimem" if fb.multicast_ptr else "p2p"
    gf = torch.cuda.CUDAGraph()
    sf = torch.cuda.Stream()
    with torch.cuda.stream(sf):
        fb.all_reduce(); torch.cuda.synchronize()
        with torch.cuda.graph(gf, stream=sf):
            for _ in range(10): fb.all_reduce()
        res["fused_unpack_" + fb.mode + "_graph"] = timed(gf.replay, 50) / 10
except Exception as e:
    res["fused_error"] = str(e)[:200]
g = torch.cuda.CUDAGraph()
try:
    sb.mode = modes[0]
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        sb.all_reduce(); torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(10): sb.all_reduce()
        res[modes[0] + "_graph"] = timed(g.replay, 50) / 10
except Exception as e:
    res["graph_error"] = str(e)[:200]
if rank == 0:
    print("allreduce us per call, %d floats (%.1f MB), world %d:" % (n, n * 4 / 1e6, world), {k: (round(v, 1) if isinstance(v, float) else v) for k, v in res.items()})
dist.destroy_process_group()
