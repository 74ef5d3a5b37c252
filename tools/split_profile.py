"""Per-kernel event timings of the split pipeline (prepare | raster | shade | backward) on config 2."""
import ctypes, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from latent_nerf_test_b200 import _lib
from bench import DeviceStep, WORKLOADS, cameras_for, load_scene, make_views
dev = torch.device("cuda:0")
w = WORKLOADS["c2"]
verts, faces, uv = load_scene(w)
geom = (verts.to(dev).float().contiguous(), faces.to(dev, torch.int32).contiguous(), uv.to(dev).float().reshape(-1, 3, 2).contiguous())
sets = []
for s in range(4):
    radius, theta, phi = make_views(w["B"], s)
    sets.append(DeviceStep(geom, w, cameras_for(radius, theta, phi, w["dy"]), 10 * s + 1, dev))
L = _lib.lib()
stream = torch.cuda.current_stream()
h = ctypes.c_void_p(stream.cuda_stream)
def step(st):
    st.prepare(h, True)
    st.shade_backward(h, stream, True)
for i in range(8): step(sets[i % 4])
torch.cuda.synchronize()
L.lp_timing_enable(1)
for i in range(40): step(sets[i % 4])
torch.cuda.synchronize()
t = _lib.collect_timings(); L.lp_timing_enable(0)
print(json.dumps({"split_us": {k: round(1e3 * v[0] / v[1], 1) for k, v in t.items()}}))
