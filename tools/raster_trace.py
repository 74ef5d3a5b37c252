"""Per-warp trace of the footprint kernel on config 2 (needs a -DLP_PROFILE library: tools/build_profile.sh, then
LP_B200_LIB=latent-nerf-test_b200/liblp_b200_profile.so python tools/raster_trace.py [ctas_per_sm]).
Prints where a launch's time goes: entry -> dependency wait -> first footprint -> last footprint, per-warp load
balance, and clocks per footprint against its candidate count."""
import ctypes, os, sys, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from latent_nerf_test_b200 import _lib

ctas = int(sys.argv[1]) if len(sys.argv) > 1 else 2
L = _lib.lib()
_lib.check(L.lp_set_option(_lib.LP_OPT_RASTER_CTAS_PER_SM, ctas))
L.lp_debug_trace.restype = ctypes.c_int
L.lp_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
dev = torch.device("cuda:0")
w = bench.WORKLOADS["c2"]
verts, faces, uv = bench.load_scene(w)
geom = (verts.to(dev).float().contiguous(), faces.to(dev, torch.int32).contiguous(), uv.to(dev).float().reshape(-1, 3, 2).contiguous())
out = {}
for s in range(2):
    st = bench.DeviceStep(geom, w, bench.workload_cameras(w, w["B"], s), 10 * s + 1, dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    for _ in range(3):
        st.prepare(stream, True)
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st.prepare(stream, False); torch.cuda.synchronize()
    e0.record(); st.raster(stream); e1.record(); torch.cuda.synchronize()
    n = 16 * 148 * 8 * 4
    buf = np.zeros(n, dtype=np.uint64)
    got = L.lp_debug_trace(buf.ctypes.data, n)
    if got <= 0:
        raise SystemExit("this library has no trace (build with tools/build_profile.sh and set LP_B200_LIB)")
    nw = 148 * ctas * 8
    t = buf[:nw * 16].reshape(nw, 16).astype(np.int64)
    t0 = t[:, 0].min()
    enter, wait, end = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3, (t[:, 2] - t0) / 1e3
    items, cands, mx, mxn, summ = t[:, 3], t[:, 4], t[:, 5], t[:, 6], t[:, 7]
    clk = 1.965e3     # clocks per us
    res = dict(set=s, ctas=ctas, warps=nw, event_us=round(e0.elapsed_time(e1) * 1e3, 1),
               enter_us=[round(float(x), 1) for x in (enter.min(), np.median(enter), enter.max())],
               wait_done_us=[round(float(x), 1) for x in (wait.min(), np.median(wait), wait.max())],
               end_us=[round(float(x), 1) for x in (end.min(), np.percentile(end, 10), np.median(end), np.percentile(end, 90), end.max())],
               items=[int(items.min()), float(np.median(items)), int(items.max()), int(items.sum())],
               cands_total=int(cands.sum()),
               busy_us=[round(float(x), 1) for x in (summ.min() / clk, np.median(summ) / clk, summ.max() / clk)],
               busy_frac=round(float(summ.sum() / clk / ((end - wait).sum() + 1e-9)), 3),
               longest_item_us=round(float(mx.max() / clk), 2), longest_item_cands=int(mxn[mx.argmax()]),
               clocks_per_item=round(float(summ.sum() / max(items.sum(), 1)), 0),
               clocks_per_cand=round(float(summ.sum() / max(cands.sum(), 1)), 0),
               frac_of_item_clocks=dict(stage=round(float(t[:, 8].sum() / summ.sum()), 3), drain=round(float(t[:, 9].sum() / summ.sum()), 3),
                                        shade=round(float(t[:, 10].sum() / summ.sum()), 3)),
               drain_rounds_per_item=round(float(t[:, 11].sum() / max(items.sum(), 1)), 2),
               clocks_per_drain_round=round(float(t[:, 9].sum() / max(t[:, 11].sum(), 1)), 0))
    print(json.dumps(res))
