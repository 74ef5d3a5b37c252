import json, os, sys, ctypes
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import latent_nerf_test_b200 as lp
from latent_nerf_test_b200 import _lib
from tests.common import scene
from bench import DeviceStep, cameras_for, make_views
DEV = "cuda:0"
sub = int(sys.argv[1]) if len(sys.argv) > 1 else 5
verts, faces, uv = scene("sphere", 0.6, 0.25, subdivide=sub)
w = dict(B=16, H=1024, W=1024, C=3, T=1024, interp="bilinear")
geom = (verts.to(DEV).float().contiguous(), faces.to(DEV, torch.int32).contiguous(), uv.to(DEV).float().reshape(-1, 3, 2).contiguous())
radius, theta, phi = make_views(16, 0)
st = DeviceStep(geom, w, cameras_for(radius, theta, phi, 0.25), 1, torch.device(DEV))
for _ in range(2): st.run()
torch.cuda.synchronize()
L = _lib.lib(); L.lp_timing_enable(1)
for _ in range(3): st.run()
torch.cuda.synchronize()
t = _lib.collect_timings(); L.lp_timing_enable(0)
print(json.dumps({"faces": int(faces.shape[0]), "us": {k: round(1e3 * v[0] / v[1], 1) for k, v in t.items()},
                  "covered_tiles": float(st.tile_any.float().mean())}))
