"""CPU count model of the tile kernel's inner loop on the bench scene (config 2): how many (footprint, face) pairs a
warp-footprint shape produces, how many survive a triangle-vs-footprint test, and the lane efficiency of the
pixel-parallel pre-test (covered pixel-face pairs / (pairs x 32 lanes)).  Guides kernel-shape decisions without GPU time."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

w = bench.WORKLOADS["c2"]
verts, faces, uv = bench.load_scene(w)
radius, theta, phi = bench.make_views(8, 0)
cams = bench.cameras_for(radius, theta, phi, w["dy"])
p = 1 / np.tan(bench.FOV / 2)
H = W = 512
fa = faces.numpy()
tris = []
for b in range(8):
    M = cams[b].numpy().astype(np.float64)
    vc = np.concatenate([verts.numpy().astype(np.float64), np.ones((len(verts), 1))], 1) @ M
    X = (vc[:, 0] * p) / (-vc[:, 2]); Y = (vc[:, 1] * p) / (-vc[:, 2])
    px = (X * W + W - 1) / 2; py = (H - 1 - Y * H) / 2          # continuous pixel coordinates (pixel centres at integers)
    tris.append(np.stack([px[fa], py[fa]], -1))
T = np.concatenate(tris)                                         # (N,3,2)
x0 = np.ceil(T[:, :, 0].min(1)).clip(0, W); x1 = np.floor(T[:, :, 0].max(1)).clip(-1, W - 1)
y0 = np.ceil(T[:, :, 1].min(1)).clip(0, H); y1 = np.floor(T[:, :, 1].max(1)).clip(-1, H - 1)
ok = (x0 <= x1) & (y0 <= y1)
T, x0, x1, y0, y1 = T[ok], x0[ok].astype(int), x1[ok].astype(int), y0[ok].astype(int), y1[ok].astype(int)
area2 = (T[:, 1, 0] - T[:, 0, 0]) * (T[:, 2, 1] - T[:, 0, 1]) - (T[:, 2, 0] - T[:, 0, 0]) * (T[:, 1, 1] - T[:, 0, 1])
covered = np.abs(area2).sum() / 2
print(f"binned faces {len(T)}, box pixels {((x1 - x0 + 1) * (y1 - y0 + 1)).sum()}, covered (pixel, face) pairs ~{covered:.0f}")

def edge_coeffs(T):
    s = np.sign(area2)[:, None]
    a = T[:, [1, 2, 0], :]; b = T[:, [2, 0, 1], :]
    A = (a[:, :, 1] - b[:, :, 1]) * s; B = (b[:, :, 0] - a[:, :, 0]) * s
    C = (a[:, :, 0] * b[:, :, 1] - a[:, :, 1] * b[:, :, 0]) * s
    return A, B, C
A, B, C = edge_coeffs(T)
for fw, fh in ((8, 4), (4, 8), (16, 2), (8, 8), (16, 4), (4, 4), (8, 2)):
    fx0, fx1, fy0, fy1 = x0 // fw, x1 // fw, y0 // fh, y1 // fh
    box_pairs = ((fx1 - fx0 + 1) * (fy1 - fy0 + 1)).sum()
    # triangle-vs-footprint: footprint survives if for every edge the best corner is inside (conservative, no margin)
    tri_pairs = 0
    nx, ny = (fx1 - fx0 + 1), (fy1 - fy0 + 1)
    small = (nx * ny) <= 64
    for i in np.nonzero(small)[0]:
        gx = (np.arange(fx0[i], fx1[i] + 1) * fw)[None, :]; gy = (np.arange(fy0[i], fy1[i] + 1) * fh)[:, None]
        keep = np.ones((ny[i], nx[i]), bool)
        for k in range(3):
            bx = gx + (fw - 1 if A[i, k] > 0 else 0); by = gy + (fh - 1 if B[i, k] > 0 else 0)
            keep &= (A[i, k] * bx + B[i, k] * by + C[i, k]) >= 0
        tri_pairs += keep.sum()
    tri_pairs += (nx * ny)[~small].sum() * 0.5                  # big faces: roughly half of the box's footprints
    lanes = fw * fh
    print(f"footprint {fw:2d}x{fh}: box pairs {box_pairs:7d}  after triangle test ~{tri_pairs:8.0f}  "
          f"lane efficiency {covered / (tri_pairs * lanes):.3f}  pair-lanes {tri_pairs * lanes / 1e6:.2f} M")
