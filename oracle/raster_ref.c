/*
 * TEST INFRASTRUCTURE — CPU oracle for the rasterizer of the Latent-Paint render path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load this.  The product (latent-nerf-test_b200/) never does.
 *
 * Restates `kaolin.render.mesh.rasterize(..., backend='cuda')` as the reference calls it
 * (reference src/latent_paint/models/render.py:42,59; src/latent_paint_mesh/models/render.py:231
 * through dibr_rasterization).  kaolin (un-pinned git master, reference setup.sh:3) is NOT
 * vendored in /root/reference and is not installable here, so this follows its documented
 * semantics with the open points fixed by decree in BASELINE.md §4 / SURVEY.md Appendix A.
 * PARITY UNPINNED at the kaolin boundary: the reference holds no test, golden vector or
 * fixture for this path.
 *
 * Arithmetic contract (fp32, evaluated left to right, no FMA contraction — build with
 * -ffp-contract=off, no -ffast-math):
 *   X,Y   = multiplier * face_vertices_image
 *   x0    = (multiplier / W) * (2 i + 1 - W)      column i, from the left
 *   y0    = (multiplier / H) * (H - 2 j - 1)      row j, from the top (+y is up)
 *   bbox  : xmin <= x0 <= xmax  and  ymin <= y0 <= ymax   (over the three scaled vertices)
 *   w0    = (Xb-x0)*(Yc-y0) - (Yb-y0)*(Xc-x0)      (w1, w2 cyclic)
 *   s     = (w0 + w1) + w2 ;  s += copysign(eps, s) ;  w_k /= s
 *   inside: w0 >= 0 and w1 >= 0 and w2 >= 0
 *   q     = (w0/za + w1/zb) + w2/zc ;  z0 = 1/q ;  rejected unless z0 < 0  (reject_behind)
 *   winner: largest z0; ties go to the lowest face index; none -> face_idx = -1
 *   w'_k  = (w_k / z_k) * z0
 *   feat_d= (w'_0 f_a,d + w'_1 f_b,d) + w'_2 f_c,d
 *
 * The open points of the decree are switches (bits of the `reject_behind` argument, which is a flag word):
 *   1  reject_behind   a face whose interpolated depth is not < 0 never wins (decree 3)
 *   2  half-open bbox  xmin <= x0 < xmax, ymin <= y0 < ymax instead of the closed box
 *   4  plain eps       s += eps instead of s += copysign(eps, s)
 *   8  affine          screen-space interpolation: z0 = (w0 za + w1 zb) + w2 zc and w'_k = w_k, instead of the
 *                      perspective-correct 1 / sum(w_k / z_k)
 * Default (1) is the decree of BASELINE.md; the others exist so that a diff against real kaolin is a flag flip.
 *
 * Two traversals that must give identical results (tests check it):
 *   lp_ref_rasterize_brute : per pixel, every face in index order (normative)
 *   lp_ref_rasterize_bbox  : per face, the pixels of its bounding box, z-buffer with the tie rule
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    float xa, ya, xb, yb, xc, yc, za, zb, zc;
    float xmin, xmax, ymin, ymax;
} face_rec;

static inline float fmin3(float a, float b, float c) { float m = a < b ? a : b; return m < c ? m : c; }
static inline float fmax3(float a, float b, float c) { float m = a > b ? a : b; return m > c ? m : c; }

static void load_face(const float *fvz, const float *fvi, float mult, face_rec *r)
{
    r->xa = mult * fvi[0]; r->ya = mult * fvi[1];
    r->xb = mult * fvi[2]; r->yb = mult * fvi[3];
    r->xc = mult * fvi[4]; r->yc = mult * fvi[5];
    r->za = fvz[0]; r->zb = fvz[1]; r->zc = fvz[2];
    r->xmin = fmin3(r->xa, r->xb, r->xc); r->xmax = fmax3(r->xa, r->xb, r->xc);
    r->ymin = fmin3(r->ya, r->yb, r->yc); r->ymax = fmax3(r->ya, r->yb, r->yc);
}

/* One (pixel, face) evaluation.  Returns 1 when the face covers the pixel and passes the
 * behind-camera rule; then *z0 and w[3] (the perspective-correct weights w') are set. */
static inline int eval_face(const face_rec *r, float x0, float y0, float eps, int flags,
                            float *z0_out, float w_out[3])
{
    const int reject_behind = flags & 1, half_open = flags & 2, plain_eps = flags & 4, affine = flags & 8;
    if (half_open) { if (!(r->xmin <= x0 && x0 < r->xmax && r->ymin <= y0 && y0 < r->ymax)) return 0; }
    else if (!(r->xmin <= x0 && x0 <= r->xmax && r->ymin <= y0 && y0 <= r->ymax)) return 0;
    float w0 = (r->xb - x0) * (r->yc - y0) - (r->yb - y0) * (r->xc - x0);
    float w1 = (r->xc - x0) * (r->ya - y0) - (r->yc - y0) * (r->xa - x0);
    float w2 = (r->xa - x0) * (r->yb - y0) - (r->ya - y0) * (r->xb - x0);
    float s = (w0 + w1) + w2;
    s = s + (plain_eps ? eps : copysignf(eps, s));
    w0 = w0 / s; w1 = w1 / s; w2 = w2 / s;
    if (!(w0 >= 0.0f && w1 >= 0.0f && w2 >= 0.0f)) return 0;
    if (affine) {
        float za0 = (w0 * r->za + w1 * r->zb) + w2 * r->zc;
        if (reject_behind) { if (!(za0 < 0.0f)) return 0; }
        else if (za0 != za0) return 0;
        *z0_out = za0;
        w_out[0] = w0; w_out[1] = w1; w_out[2] = w2;
        return 1;
    }
    float q = (w0 / r->za + w1 / r->zb) + w2 / r->zc;
    float z0 = 1.0f / q;
    if (reject_behind) { if (!(z0 < 0.0f)) return 0; }
    else if (z0 != z0) return 0;
    *z0_out = z0;
    w_out[0] = (w0 / r->za) * z0;
    w_out[1] = (w1 / r->zb) * z0;
    w_out[2] = (w2 / r->zc) * z0;
    return 1;
}

static void write_pixel(int64_t pix, int64_t bf, int D, const float *feat, const float w[3], float z0,
                        float *out_feat, int64_t *face_idx, float *out_w, float *out_depth, int f)
{
    face_idx[pix] = f;
    if (out_w) { out_w[pix * 3 + 0] = w[0]; out_w[pix * 3 + 1] = w[1]; out_w[pix * 3 + 2] = w[2]; }
    if (out_depth) out_depth[pix] = z0;
    if (out_feat && feat) {
        const float *fa = feat + (bf * 3 + 0) * D, *fb = feat + (bf * 3 + 1) * D, *fc = feat + (bf * 3 + 2) * D;
        for (int d = 0; d < D; ++d)
            out_feat[pix * D + d] = (w[0] * fa[d] + w[1] * fb[d]) + w[2] * fc[d];
    }
}

static void clear_outputs(int64_t n, int D, float *out_feat, int64_t *face_idx, float *out_w, float *out_depth)
{
    for (int64_t p = 0; p < n; ++p) face_idx[p] = -1;
    if (out_feat) memset(out_feat, 0, sizeof(float) * n * D);
    if (out_w) memset(out_w, 0, sizeof(float) * n * 3);
    if (out_depth) memset(out_depth, 0, sizeof(float) * n);
}

/* fvz (B,F,3)  fvi (B,F,3,2)  feat (B,F,3,D) or NULL  valid (B,F) bytes or NULL
 * out_feat (B,H,W,D)  face_idx (B,H,W) int64  out_w (B,H,W,3) or NULL  out_depth (B,H,W) or NULL */
void lp_ref_rasterize_brute(int B, int F, int H, int W, int D,
                            const float *fvz, const float *fvi, const float *feat, const uint8_t *valid,
                            float multiplier, float eps, int reject_behind,
                            float *out_feat, int64_t *face_idx, float *out_w, float *out_depth)
{
    clear_outputs((int64_t)B * H * W, D, out_feat, face_idx, out_w, out_depth);
    face_rec *recs = (face_rec *)malloc(sizeof(face_rec) * (size_t)(F > 0 ? F : 1));
    for (int b = 0; b < B; ++b) {
        for (int f = 0; f < F; ++f) {
            int64_t bf = (int64_t)b * F + f;
            load_face(fvz + bf * 3, fvi + bf * 6, multiplier, &recs[f]);
        }
        #pragma omp parallel for schedule(dynamic, 4)
        for (int j = 0; j < H; ++j) {
            float y0 = (multiplier / (float)H) * (float)(H - 2 * j - 1);
            for (int i = 0; i < W; ++i) {
                float x0 = (multiplier / (float)W) * (float)(2 * i + 1 - W);
                float best = -INFINITY, bw[3] = {0, 0, 0};
                int bestf = -1;
                for (int f = 0; f < F; ++f) {
                    if (valid && !valid[(int64_t)b * F + f]) continue;
                    float z0, w[3];
                    if (!eval_face(&recs[f], x0, y0, eps, reject_behind, &z0, w)) continue;
                    if (bestf < 0 || z0 > best) { best = z0; bestf = f; bw[0] = w[0]; bw[1] = w[1]; bw[2] = w[2]; }
                }
                if (bestf >= 0) {
                    int64_t pix = ((int64_t)b * H + j) * W + i;
                    write_pixel(pix, (int64_t)b * F + bestf, D, feat, bw, best, out_feat, face_idx, out_w, out_depth, bestf);
                }
            }
        }
    }
    free(recs);
}

/* Smallest i in [0,n] with coord(i) >= v (or coord decreasing variant handled by caller). */
static int first_col_ge(float v, int W, float mult)
{
    /* x0(i) is non-decreasing in i: plain binary search on the exact expression */
    int lo = 0, hi = W;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        float x0 = (mult / (float)W) * (float)(2 * mid + 1 - W);
        if (x0 >= v) hi = mid; else lo = mid + 1;
    }
    return lo;
}
static int last_col_le(float v, int W, float mult)
{
    int lo = -1, hi = W - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        float x0 = (mult / (float)W) * (float)(2 * mid + 1 - W);
        if (x0 <= v) lo = mid; else hi = mid - 1;
    }
    return lo;
}
/* y0(j) is non-increasing in j */
static int first_row_le(float v, int H, float mult)
{
    int lo = 0, hi = H;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        float y0 = (mult / (float)H) * (float)(H - 2 * mid - 1);
        if (y0 <= v) hi = mid; else lo = mid + 1;
    }
    return lo;
}
static int last_row_ge(float v, int H, float mult)
{
    int lo = -1, hi = H - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        float y0 = (mult / (float)H) * (float)(H - 2 * mid - 1);
        if (y0 >= v) lo = mid; else hi = mid - 1;
    }
    return lo;
}

void lp_ref_rasterize_bbox(int B, int F, int H, int W, int D,
                           const float *fvz, const float *fvi, const float *feat, const uint8_t *valid,
                           float multiplier, float eps, int reject_behind,
                           float *out_feat, int64_t *face_idx, float *out_w, float *out_depth)
{
    int64_t npix = (int64_t)H * W;
    clear_outputs((int64_t)B * npix, D, out_feat, face_idx, out_w, out_depth);
    #pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        float *zbuf = (float *)malloc(sizeof(float) * npix);
        float *wbuf = (float *)malloc(sizeof(float) * npix * 3);
        int64_t *fi = face_idx + (int64_t)b * npix;
        for (int f = 0; f < F; ++f) {
            int64_t bf = (int64_t)b * F + f;
            if (valid && !valid[bf]) continue;
            face_rec r;
            load_face(fvz + bf * 3, fvi + bf * 6, multiplier, &r);
            if (!(r.xmin <= r.xmax) || !(r.ymin <= r.ymax)) continue; /* NaN boxes never pass the bbox test */
            int i0 = first_col_ge(r.xmin, W, multiplier), i1 = last_col_le(r.xmax, W, multiplier);
            int j0 = first_row_le(r.ymax, H, multiplier), j1 = last_row_ge(r.ymin, H, multiplier);
            for (int j = j0; j <= j1; ++j) {
                float y0 = (multiplier / (float)H) * (float)(H - 2 * j - 1);
                for (int i = i0; i <= i1; ++i) {
                    float x0 = (multiplier / (float)W) * (float)(2 * i + 1 - W);
                    float z0, w[3];
                    if (!eval_face(&r, x0, y0, eps, reject_behind, &z0, w)) continue;
                    int64_t p = (int64_t)j * W + i;
                    /* faces arrive in increasing index, so strict > keeps the lowest index on ties */
                    if (fi[p] < 0 || z0 > zbuf[p]) {
                        fi[p] = f; zbuf[p] = z0;
                        wbuf[p * 3] = w[0]; wbuf[p * 3 + 1] = w[1]; wbuf[p * 3 + 2] = w[2];
                    }
                }
            }
        }
        for (int64_t p = 0; p < npix; ++p)
            if (fi[p] >= 0) {
                int f = (int)fi[p];
                write_pixel((int64_t)b * npix + p, (int64_t)b * F + f, D, feat, wbuf + p * 3, zbuf[p],
                            out_feat, face_idx, out_w, out_depth, f);
            }
        free(zbuf); free(wbuf);
    }
}

int lp_ref_version(void) { return 1; }
