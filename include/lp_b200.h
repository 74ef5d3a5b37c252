/*
 * lp_b200 — C ABI of the B200-native Latent-Paint mesh renderer (sm_100a).
 *
 * Drop-in boundary: these entry points are what the reference's renderer would bind in place
 * of the kaolin / ATen calls on its render path.  Each one names the reference interface it
 * replaces (paths relative to the reference checkout):
 *
 *   lp_cameras_from_views   Renderer.get_camera_from_view            src/latent_paint/models/render.py:19-31
 *                                                                    src/latent_paint_mesh/models/render.py:42-55
 *                           (+ kal.render.camera.generate_transformation_matrix)
 *   lp_render_forward       kal.render.mesh.prepare_vertices         src/latent_paint/models/render.py:39,56
 *                           kal.render.mesh.rasterize                src/latent_paint/models/render.py:42,59
 *                           kal.render.mesh.dibr_rasterization       src/latent_paint_mesh/models/render.py:231
 *                           kal.render.mesh.texture_mapping          src/latent_paint/models/render.py:64
 *                                                                    src/latent_paint_mesh/models/render.py:243
 *                           mask / white background composition      src/latent_paint/models/render.py:45,63-67
 *                           kal.render.mesh.spherical_harmonic_lighting  src/latent_paint_mesh/models/render.py:258-259
 *   lp_vertex_normals       Renderer.compute_vertex_normals          src/latent_paint_mesh/models/render.py:57-105
 *                           (+ kal.ops.mesh.index_vertices_by_faces, :202 — gathered inside lp_render_forward)
 *   lp_render_backward      autograd backward of grid_sample (texture_mapping) and of kaolin's rasterize
 *                           into the face features: pred.backward(gradient=grad)
 *                                                                    src/latent_paint_mesh/training/trainer.py:656-660
 *   lp_render_step_host     one forward+backward through host buffers (what bench.py's e2e times)
 *   lp_render_forward with under_image / under_mask / composed
 *                           pred_back * (1 - mask) + pred_features * mask   src/latent_paint/models/textured_mesh.py:211-212
 *   lp_allreduce_*          the texture-gradient sum over the ranks (new: the reference is single-GPU)
 *   lp_adam_step            torch.optim.Adam(lr, betas=(0.9, 0.99), eps=1e-15).step()
 *                                                                    src/latent_paint/training/trainer.py:93-95
 *                                                                    src/latent_paint_mesh/training/trainer.py:326-328
 *
 * Conventions: every pointer is caller-owned; device pointers unless the name ends in _host.
 * All floating point is fp32, indices int32.  Calls only enqueue work on the given stream
 * (a cudaStream_t passed as void*); they never synchronise the device (lp_render_step_host
 * excepted) and allocate nothing.  Return value: LP_OK or an LP_ERR_* code;
 * lp_last_error() gives the message of the last failure on the calling thread.
 */
#ifndef LP_B200_H
#define LP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LP_B200_VERSION 100

enum {
    LP_OK = 0,
    LP_ERR_BAD_ARG = 1,      /* null pointer, non-positive size, inconsistent shapes */
    LP_ERR_UNSUPPORTED = 2,  /* e.g. too many channels, sizes beyond the index range */
    LP_ERR_WORKSPACE = 3,    /* workspace missing or smaller than lp_workspace_bytes() */
    LP_ERR_CUDA = 4          /* a CUDA runtime call failed; see lp_last_error() */
};

enum { LP_INTERP_NEAREST = 0, LP_INTERP_BILINEAR = 1, LP_INTERP_BICUBIC = 2 };   /* F.grid_sample modes the reference accepts (render.py:9) */

/* LpForwardArgs.flags / LpBackwardArgs.flags */
enum {
    LP_FLAG_MASK_IMAGE       = 1u << 0, /* image *= (face_idx > -1); mask output is that 0/1 mask
                                           (latent_paint flavour, render.py:63-65).  Without it the
                                           image is not masked and the mask output is the interpolated
                                           all-ones feature (latent_paint_mesh flavour, render.py:224-243) */
    LP_FLAG_WHITE_BACKGROUND = 1u << 1, /* image += 1 * (1 - mask) */
    LP_FLAG_REJECT_BEHIND    = 1u << 2, /* faces whose interpolated depth is not < 0 never win (BASELINE.md decree 3) */
    LP_FLAG_CULL_NZ_ZERO     = 1u << 3, /* drop faces whose unit camera-space normal has |n_z| == 0:
                                           dibr_rasterization's valid_faces as the reference calls it with abs() */
    LP_FLAG_SHADE_FEATURES   = 1u << 4, /* interpolate face_features instead of sampling a texture
                                           (Renderer.render_single_view, latent_paint render.py:34-47) */
    LP_FLAG_GRAD_OVERWRITE   = 1u << 5, /* lp_render_backward with a workspace: grad_texture is written, not
                                           accumulated into, so the caller need not zero it */
    LP_FLAG_GRAD_INTERLEAVED = 1u << 6, /* lp_render_backward with a workspace: leave the gradient in the workspace as
                                           (Th,Tw,4) texel-interleaved float4 and do not touch grad_texture —
                                           lp_allreduce_unpack sums it over the ranks and writes the planar gradient */
    /* The open points of the kaolin restatement as switches; 0 = the decree.  The same
       switches exist in oracle/raster_ref.c and oracle/kaolin_shim.py, so a diff against real kaolin is a flag flip */
    LP_FLAG_BBOX_HALF_OPEN   = 1u << 7,  /* bounding-box test xmin <= x0 < xmax, ymin <= y0 < ymax instead of the closed box */
    LP_FLAG_PLAIN_EPS        = 1u << 8,  /* s += eps instead of s += copysign(eps, s) */
    LP_FLAG_AFFINE_INTERP    = 1u << 9,  /* screen-space interpolation: z0 = sum w_k z_k, w'_k = w_k, instead of perspective-correct */
    LP_FLAG_SH_BAND1_XZY     = 1u << 10, /* SH band-1 axis order (x, z, y) instead of (y, z, x) */
    LP_FLAG_GRAD_NO_CLEAR    = 1u << 11, /* lp_render_backward with a workspace: the caller has already zeroed the first
                                            Th * Tw * 16 bytes of the workspace (e.g. on another stream, next to the texture
                                            fetch), so the call does not clear it again */
    LP_FLAG_MICRO_OFF        = 1u << 22, /* never / always rasterize small faces face-parallel in the setup kernel */
    LP_FLAG_MICRO_ON         = 1u << 23  /* (default: on when the mesh has at least one face per 16 pixels) */
    /* bits 24-30: stop-after-stage ablation switches, compiled in only with -DLP_PROFILE (tools/), ignored otherwise */
};

typedef struct LpForwardArgs {
    /* geometry — prepare_vertices inputs */
    const float   *verts;          /* (V,3) */
    const int32_t *faces;          /* (F,3) */
    int32_t        V, F;
    /* alternative geometry input = kal.render.mesh.rasterize's own arguments (verts/faces/cameras unused):
       vertices already projected by the caller */
    const float   *face_vertices_image; /* (B,F,3,2) NDC xy, or NULL for the vertex path */
    const float   *face_vertices_z;     /* (B,F,3) camera-space z */
    const uint8_t *valid_faces;         /* (B,F) or NULL: 0 drops the face (dibr_rasterization's back-face rule) */
    /* views */
    const float   *cameras;        /* (B,4,3) look-at matrices [R;t]: v_cam = [v,1] @ M */
    int32_t        B;
    float          proj[3];        /* camera_proj vector: (1/tan(fov/2), 1/tan(fov/2), -1) */
    int32_t        H, W;           /* output rows, columns  (the reference passes dims[1], dims[0]) */
    float          multiplier;     /* kaolin rasterize multiplier (1000) */
    float          eps;            /* kaolin rasterize eps (1e-8) */
    uint32_t       flags;
    /* texture shading (flags without LP_FLAG_SHADE_FEATURES) */
    const float   *face_uv;        /* (F,3,2) per-face-corner UVs, shared by all views */
    const float   *texture;        /* (C,Th,Tw) planar, the reference's (1,C,T,T) parameter */
    int32_t        C, Th, Tw;
    int32_t        interp;         /* LP_INTERP_* */
    /* face-feature shading (LP_FLAG_SHADE_FEATURES) */
    const float   *face_features;  /* (Bf,F,3,D) with Bf = 1 or B */
    int32_t        D, features_batched;
    /* latent_paint_mesh extras: needed only when the normals / lighting outputs are requested.
       The call then also runs the lp_vertex_normals step between face setup and shading. */
    const int32_t *vf_offsets;     /* (V+1) vertex -> incident-corner CSR, see lp_vertex_normals */
    const int32_t *vf_faces;       /* (3F) */
    float         *face_normals;   /* out/scratch (B,F,3): unit camera-space face normals */
    float         *vertex_normals; /* out/scratch (B,V,3): averaged, not re-normalised */
    const float   *lights;         /* (9) SH coefficients */
    /* outputs; image and mask are required, the rest may be NULL */
    float         *image;          /* (B,C|D,H,W) */
    float         *mask;           /* (B,1,H,W) */
    float         *uv;             /* (B,H,W,2) interpolated UVs, saved for lp_render_backward;
                                      with LP_FLAG_MASK_IMAGE uncovered pixels hold u = NaN (coordinates may be negative) */
    int32_t       *face_idx;       /* (B,H,W) winning face, -1 = none */
    float         *bary;           /* (B,H,W,3) perspective-correct weights w' */
    float         *depth;          /* (B,H,W) camera-space z of the visible surface, 0 = none */
    float         *normals;        /* (B,3,H,W) interpolated averaged vertex normals */
    float         *lighting;       /* (B,1,H,W) clamp(SH(normals)·lights, 1e-8, 1) */
    uint8_t       *footprint_any;  /* optional (B, ceil(H/4), ceil(W/8)): 1 where the 8 x 4-pixel footprint (the pixels one warp
                                      rasterizes) holds a covered pixel.  With LP_FLAG_MASK_IMAGE the saved uv of the other
                                      footprints is then NOT written, so the same buffer must be passed to lp_render_backward */
    /* scratch */
    void          *workspace;
    uint64_t       workspace_bytes;
    /* model-level composition fused into the face-feature shading (LP_FLAG_SHADE_FEATURES only; all three or none):
       composed = image * (1 - under_mask) + under_image * under_mask, the reference's
       pred_back * (1 - mask) + pred_features * mask (src/latent_paint/models/textured_mesh.py:211-212) with this
       call rendering pred_back (the env sphere) over an earlier texture render (under_image, under_mask) */
    const float   *under_image;    /* (B,D,H,W) */
    const float   *under_mask;     /* (B,1,H,W) */
    float         *composed;       /* out (B,D,H,W) */
    /* optional: the texture once more as (Th,Tw,4) texel-interleaved float4 (lp_pack_texture; C <= 4).  lp_render_shade
       then fetches a tap with one 16-byte load instead of C loads Th*Tw apart; results are the same bits */
    const void    *texture_rgba;
} LpForwardArgs;

typedef struct LpBackwardArgs {
    int32_t        B, H, W;
    uint32_t       flags;          /* same flags as the forward call */
    const float   *grad_image;     /* (B,C|D,H,W) dL/d image */
    /* texture path */
    const float   *uv;             /* (B,H,W,2) saved by the forward */
    int32_t        C, Th, Tw, interp;
    float         *grad_texture;   /* (C,Th,Tw) planar; ACCUMULATED into (caller zeroes it) */
    int64_t        grad_texture_batch_stride; /* elements between the textures of consecutive views; 0 = one
                                      texture shared by all views (the kaolin-level texture_mapping takes (B,C,T,T)) */
    /* face-feature path */
    const int32_t *face_idx;       /* (B,H,W) */
    const float   *bary;           /* (B,H,W,3) */
    int32_t        F, D, features_batched;
    float         *grad_face_features; /* (Bf,F,3,D); ACCUMULATED into */
    const uint8_t *footprint_any;       /* optional, written by the forward call (used with LP_FLAG_MASK_IMAGE) */
    /* optional scratch of lp_backward_workspace_bytes(): with it (and C <= 4) the taps are accumulated with
       16-byte vector REDs into a texel-interleaved (Th,Tw,4) buffer and then unpacked into grad_texture —
       a third of the atomic operations of the planar path */
    void          *workspace;
    uint64_t       workspace_bytes;
    const float   *under_mask;     /* optional (B,1,H,W), face-feature path: grad_image is dL/d composed and is scaled
                                      by (1 - under_mask) per pixel (the backward of the fused composition) */
    /* optional (both or none; LP_FLAG_MASK_IMAGE): the forward call's list of covered footprints (8 x 4 pixels that hold
       at least one covered pixel, written by the footprint kernel) and the control block holding its length, from
       lp_forward_worklist(); valid while that call's workspace has not been reused.  The backward then visits the
       listed footprints directly instead of scanning the coverage flags of all of them */
    const void    *worklist;
    const void    *worklist_ctrl;
} LpBackwardArgs;

/* kal.render.mesh.texture_mapping forward (latent_paint render.py:64, latent_paint_mesh render.py:243):
 * clamp, flip v, ATen grid_sample(align_corners=False, padding_mode='border') texel arithmetic */
typedef struct LpTextureMapArgs {
    int32_t        B, H, W;
    const float   *uv;             /* (B,H,W,2) texture coordinates */
    const float   *texture;        /* (Bt,C,Th,Tw) planar, Bt = 1 (stride 0) or B */
    int64_t        texture_batch_stride;
    int32_t        C, Th, Tw, interp;
    float         *out;            /* (B,C,H,W) */
} LpTextureMapArgs;

int         lp_version(void);
/* builds with -DLP_CHECKED test the kernels' index and capacity invariants on the device: number of violations since the
 * library was loaded (0 in ordinary builds), *first_line = source line of the first; synchronises the device */
int         lp_check_failures(int *first_line);
/* -DLP_PROFILE builds: per-warp trace of the footprint kernel's last launch, 16 words per warp (entry, after the
 * dependency wait, exit in ns of %globaltimer; footprints, candidates, longest footprint in clocks and its candidates,
 * sum of footprint clocks, clocks in staging / exact evaluation / shading, evaluation rounds); returns the number of words copied, 0 in production builds */
int         lp_debug_trace(unsigned long long *out, int words);
/* process-wide switches.  LP_OPT_PDL (default 1): chain the kernels of a call with programmatic dependent launch
 * (each kernel's prologue overlaps its predecessor's tail); 0 = plain stream-ordered launches.
 * LP_OPT_RASTER_CTAS_PER_SM (default 0 = all the tile kernel's launch bounds allow): resident CTAs per SM of the
 * persistent tile kernel; fewer leave room for kernels of other streams to run beside it */
enum { LP_OPT_PDL = 1, LP_OPT_RASTER_CTAS_PER_SM = 2, LP_OPT_EXCHANGE_CTAS = 3 /* CTAs of lp_exchange_step, 0 = one per SM */,
       LP_OPT_WALK_CTAS_PER_SM = 4 /* CTAs per SM of lp_render_shade / lp_render_backward, 0 = eight */,
       LP_OPT_EXCHANGE_BULK = 5 /* peer form of lp_exchange_step reads with bulk asynchronous copies (default 1) */ };
int         lp_set_option(int option, int value);
const char *lp_last_error(void);
const char *lp_error_string(int code);

/* bytes of device scratch lp_render_forward needs for B views of F faces at H x W */
uint64_t    lp_workspace_bytes(int32_t B, int32_t F, int32_t H, int32_t W);

/* elev/azim/radius: (B) device arrays (radius_stride 0 broadcasts one value); cameras out (B,4,3) */
int lp_cameras_from_views(const float *elev, const float *azim, const float *radius, int32_t radius_stride,
                          float look_at_height, int32_t B, float *cameras, void *stream);

uint64_t    lp_backward_workspace_bytes(int32_t C, int32_t Th, int32_t Tw);
/* where in `args->workspace` the forward leaves its covered-footprint list and its control block (for
 * LpBackwardArgs.worklist / worklist_ctrl) */
int lp_forward_worklist(const LpForwardArgs *args, const void **worklist, const void **worklist_ctrl);

int lp_render_forward(const LpForwardArgs *args, void *stream);
/* lp_render_forward split in three, so a caller can overlap everything that does not read the texture
 * (geometry, binning, visibility, uv / mask / normals) of the NEXT batch with the texture fetch, the backward and
 * the gradient all-reduce of the current one on another stream:
 *   lp_render_prepare  stage 1 + binning                      (needs verts / cameras, fills the workspace)
 *   lp_render_raster   stage 2-3: visibility, uv, mask, optional buffers  (needs the prepared workspace)
 *   lp_render_shade    stage 4: texture fetch + composition -> image     (needs uv [, mask, footprint_any] of the raster call)
 * All three take the same argument block as lp_render_forward, which runs 1-4 with the fetch fused into the tile kernel. */
int lp_render_prepare(const LpForwardArgs *args, void *stream);
int lp_render_raster(const LpForwardArgs *args, void *stream);
int lp_render_shade(const LpForwardArgs *args, void *stream);
int lp_render_raster_shade(const LpForwardArgs *args, void *stream);   /* raster + shade in the one fused tile kernel */
int lp_render_backward(const LpBackwardArgs *args, void *stream);
int lp_texture_map_forward(const LpTextureMapArgs *args, void *stream);
/* The resize to the latent grid that follows the render in TexturedMeshModel.render_train (src/latent_paint/models/
 * textured_mesh.py:214-218: F.interpolate(x, (64, 64), mode='bicubic') on mask, background, foreground and image) as one
 * launch over up to eight (planes, H, W) -> (planes, OH, OW) tensors, with ATen's upsample_bicubic2d arithmetic
 * (align_corners=False).  backward != 0: `in` holds the upstream gradients (planes, OH, OW), `out` the gradients at
 * (planes, H, W), ACCUMULATED into (the caller zeroes them). */
typedef struct LpResizeArgs {
    const float *in[8];
    float       *out[8];
    int32_t      planes[8];        /* B * C of each tensor */
    int32_t      n, H, W, OH, OW, backward;
} LpResizeArgs;
int lp_resize_bicubic(const LpResizeArgs *args, void *stream);

/* planar (C,Th,Tw) texture -> (Th,Tw,4) texel-interleaved float4 for LpForwardArgs.texture_rgba (16 * Th * Tw bytes);
 * repack whenever the texture changes (after the optimiser step) */
int lp_pack_texture(const float *texture, int32_t C, int32_t Th, int32_t Tw, void *texture_rgba, void *stream);

/* vertex → incident (corner-major, face-ascending) CSR: offsets (V+1), entries (3F) hold face ids.
 * face_normals (B,F,3) → vertex_normals (B,V,3) = mean of incident unit face normals, not re-normalised */
int lp_vertex_normals(const float *face_normals, const int32_t *vf_offsets, const int32_t *vf_faces,
                      int32_t B, int32_t V, int32_t F, float *vertex_normals, void *stream);

/* One latent_paint-flavour forward + backward through HOST buffers: copies cameras and grad_image
 * host→device, renders, scatters the gradient, copies image, mask and grad_texture device→host and
 * synchronises the stream.  Geometry, texture and all device scratch are caller-provided device
 * memory referenced by `fwd` / `bwd` (their image/mask/uv/grad pointers are device staging). */
int lp_render_step_host(const LpForwardArgs *fwd, const LpBackwardArgs *bwd,
                        const float *cameras_host, const float *grad_image_host,
                        float *image_host, float *mask_host, float *grad_texture_host, void *stream);

/* The path's one exchange step (SURVEY.md §8 e): sum the flat texture-gradient buffer over the ranks of one box.
 * The buffer of every rank must live in symmetric memory (same size, mapped into every peer); the caller puts a
 * stream-ordered barrier over all ranks BEFORE (every rank's backward finished) and AFTER each call (and between
 * the two phases of the p2p form).  count = number of floats, a multiple of 4.
 *   lp_allreduce_multimem  NVSwitch in-switch reduction (multimem.ld_reduce / multimem.st on the multicast address)
 *   lp_allreduce_p2p       phase 0 reduce-scatter, phase 1 all-gather over plain peer pointers (device array of
 *                          `world` buffer pointers) — fallback when the box has no multicast support */
int lp_allreduce_multimem(void *multicast_ptr, int64_t count, int32_t rank, int32_t world, void *stream);
int lp_allreduce_p2p(void *const *buffer_ptrs_dev, int64_t count, int32_t rank, int32_t world, int32_t phase, void *stream);

/* The step after the backward (SURVEY.md §8 f rank 4): torch.optim.Adam on the texture — the reference's optimiser,
 * Adam(lr, betas=(0.9, 0.99), eps=1e-15), src/latent_paint/training/trainer.py:93-95 — as ONE kernel, optionally
 * fused with the unpack of the vector-RED backward.  The gradient comes either planar (`grad`, (C,ntex)) or
 * texel-interleaved (`accum`, (ntex,4) float4, what lp_render_backward leaves with LP_FLAG_GRAD_INTERLEAVED; then
 * C <= 4 and, if `grad` is non-NULL, the planar gradient is written there as well).  param / exp_avg / exp_avg_sq
 * are planar (C,ntex) and updated in place with torch's single-tensor Adam arithmetic (no weight decay, no
 * amsgrad): m += (g - m)(1 - b1); v = v b2 + (1 - b2) g g; p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps).
 * `step` is the 1-based step count AFTER this update. */
typedef struct LpAdamArgs {
    const void  *accum;            /* (ntex,4) float4 or NULL */
    float       *grad;             /* (C,ntex): input when accum is NULL, optional output otherwise */
    float       *param, *exp_avg, *exp_avg_sq;   /* (C,ntex) */
    int64_t      ntex;
    int32_t      C;
    float        lr, beta1, beta2, eps;
    int32_t      step;
} LpAdamArgs;
int lp_adam_step(const LpAdamArgs *args, void *stream);

/* Exchange fused with the unpack of the vector-RED backward: every rank holds one symmetric allocation with the
 * texel-interleaved accumulation buffer (ntex float4, what lp_render_backward leaves with LP_FLAG_GRAD_INTERLEAVED)
 * at byte offset accum_offset and the planar (C,ntex) gradient at byte offset grad_offset.  Rank r sums slice r of
 * the accumulation buffers of all ranks (multicast_base != NULL: multimem.ld_reduce inside the NVSwitch; else peer
 * loads in rank order), transposes it to planar in registers and writes it into the gradient of EVERY rank
 * (multimem.st / peer stores): one kernel instead of unpack + reduce-scatter + all-gather, and every rank ends with
 * bit-identical sums.  ntex must be a multiple of 4 * world.  Barriers before and after are the caller's, as above. */
int lp_allreduce_unpack(void *multicast_base, void *const *buffer_ptrs_dev, uint64_t accum_offset, uint64_t grad_offset,
                        int64_t ntex, int32_t C, int32_t rank, int32_t world, void *stream);

/* The exchange as ONE launch, handshakes included: every rank's symmetric allocation also holds LP_EXCHANGE_FLAG_BYTES of
 * flag words at flags_offset (zeroed once by the caller).  Inside the kernel rank r tells every peer that its backward is
 * complete, waits for theirs, reduces + unpacks + broadcasts slice r as lp_allreduce_unpack does, and leaves only when every
 * peer's broadcast has landed — no host-enqueued barrier before, between or after.  Replayable from CUDA graphs (the epoch
 * lives in the flag block).  With adam != 0 the broadcast is the UPDATED PARAMETER slice instead of the gradient: rank r
 * owns slice r of the optimiser state (exp_avg / exp_avg_sq: local, planar (C, ntex / world)), applies torch's Adam step
 * to slice r of the planar (C, ntex) parameters at param_offset and stores it to every rank (reduce slice -> Adam on the
 * slice -> broadcast parameters; src/latent_paint_mesh/training/trainer.py:326-328 + :407 sharded over the ranks). */
#define LP_EXCHANGE_FLAG_BYTES 8192
typedef struct LpExchangeArgs {
    void        *multicast_base;   /* multicast mapping of the allocation, or NULL: peer loads / stores */
    void *const *buffer_ptrs_dev;  /* device array of `world` allocation base pointers */
    uint64_t     accum_offset, grad_offset, flags_offset;   /* bytes from the allocation base */
    int64_t      ntex;             /* texels, a multiple of 4 * world */
    int32_t      C, rank, world;
    int32_t      adam;             /* 0: broadcast the summed gradient to grad_offset; 1: optimiser epilogue */
    uint64_t     param_offset;     /* adam: planar (C, ntex) parameters in the symmetric allocation */
    float       *exp_avg, *exp_avg_sq;   /* adam: this rank's slices, planar (C, ntex / world) */
    float        lr, beta1, beta2, eps;
    int32_t      step;             /* adam: 1-based step count after this update */
} LpExchangeArgs;
int lp_exchange_step(const LpExchangeArgs *args, void *stream);

/* Instrumentation (bench.py's roofline leg): while enabled, every kernel launch of this library is
 * bracketed by CUDA events on its stream.  lp_timing_collect waits for them, sums the elapsed
 * milliseconds per kernel name into total_ms[]/counts[] (names[] receives static strings), clears
 * the record and returns the number of distinct names, or -LP_ERR_CUDA.  Not
 * usable during stream capture. */
int lp_timing_enable(int on);
int lp_timing_collect(int max_names, const char **names, float *total_ms, int *counts);

/* Same, without the final synchronisation: with pinned host buffers the copies are asynchronous, so steps issued
 * on different streams overlap their host->device copy, kernels and device->host copy (full-duplex PCIe).  The host
 * results are valid once the stream has been synchronised. */
int lp_render_step_host_async(const LpForwardArgs *fwd, const LpBackwardArgs *bwd,
                              const float *cameras_host, const float *grad_image_host,
                              float *image_host, float *mask_host, float *grad_texture_host, void *stream);

/* number of kernel launches the last lp_render_forward / lp_render_backward on this thread enqueued */
int lp_last_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* LP_B200_H */
