"""Where do the device-side invariant violations on config 4 come from?  Needs the -DLP_CHECKED library
(LP_B200_LIB=latent-nerf-test_b200/liblp_b200_checked.so).  Runs the split forward + backward of one buffer set eagerly,
from a CUDA graph, and two sets on two streams, and prints the violation counter after each."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from latent_nerf_test_b200 import _lib

L = _lib.lib()
dev = torch.device("cuda:0")
w = dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c4"])
if len(sys.argv) > 2:
    w["B"] = int(sys.argv[2])
verts, faces, uv = bench.load_scene(w)
geom = (verts.to(dev).float().contiguous(), faces.to(dev, torch.int32).contiguous(), uv.to(dev).float().reshape(-1, 3, 2).contiguous())
def bad(tag):
    torch.cuda.synchronize()
    line = ctypes.c_int32(0)
    print(tag, "violations so far:", L.lp_check_failures(ctypes.byref(line)), "first line", line.value, flush=True)
sets = [bench.DeviceStep(geom, w, bench.workload_cameras(w, w["B"], s), 10 * s + 1, dev) for s in range(2)]
s0 = torch.cuda.Stream(dev); s1 = torch.cuda.Stream(dev)
h0, h1 = ctypes.c_void_p(s0.cuda_stream), ctypes.c_void_p(s1.cuda_stream)
with torch.cuda.stream(s0):
    for _ in range(3):
        sets[0].prepare(h0, True); sets[0].shade_backward(h0, s0, True)
bad("eager, one set, one stream")
for ctas in (2, 4):
    _lib.check(L.lp_set_option(_lib.LP_OPT_RASTER_CTAS_PER_SM, ctas))
    with torch.cuda.stream(s0):
        for _ in range(2):
            sets[0].prepare(h0, True); sets[0].shade_backward(h0, s0, True)
    bad(f"eager, {ctas} footprint-kernel CTAs per SM")
g = torch.cuda.CUDAGraph()
with torch.cuda.stream(s0):
    with torch.cuda.graph(g, stream=s0):
        sets[0].prepare(h0, True); sets[0].shade_backward(h0, s0, True)
    for _ in range(3):
        g.replay()
bad("graph, one set")
ev = [torch.cuda.Event() for _ in range(2)]
g2 = torch.cuda.CUDAGraph()
with torch.cuda.stream(s0):
    with torch.cuda.graph(g2, stream=s0):
        s1.wait_stream(s0)
        for i in range(4):
            k = i % 2
            sets[k].prepare(h1, True)          # geometry + raster of set k on the second stream ...
            ev[k].record(s1)
            s0.wait_event(ev[k])
            sets[k].shade_backward(h0, s0, True)   # ... its shade + backward on the first (the next prepare of set k is
            s1.wait_stream(s0)                     # ordered behind it)
        s0.wait_stream(s1)
    for _ in range(3):
        g2.replay()
bad("graph, two sets on two streams")
