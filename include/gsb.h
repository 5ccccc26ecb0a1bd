/*
 * gsb.h -- C ABI of libgsb.so: the B200-native (sm_100a) differentiable 3D-Gaussian
 * rasterizer that replaces GSORB-SLAM's libCudaRasterizer.so + simple_knn.
 *
 * Every entry point below is what the reference's C++ seam for this path binds today:
 *
 *   gsb_forward / gsb_forward_ws   <- CudaRasterizer::Rasterizer::forward
 *                                     (Thirdparty/diff_gaussian_rasterization/cuda_rasterizer/rasterizer.h:31-53,
 *                                      called from src/Rasterizer.cu:191)
 *   gsb_backward                   <- CudaRasterizer::Rasterizer::backward  (rasterizer.h:55-83, src/Rasterizer.cu:265)
 *   gsb_visible_filter             <- CudaRasterizer::Rasterizer::visible_filter (rasterizer.h:85-100, src/Rasterizer.cu:365)
 *   gsb_mark_visible               <- CudaRasterizer::Rasterizer::markVisible (rasterizer.h:24-29, src/Rasterizer.cu:310)
 *   gsb_knn_mean_dist2             <- SimpleKNN::knn (include/simple_knn.h:15-19, src/spatial.cu:24)
 *
 * plus fused extensions of the caller-side prologue/epilogue:
 *
 *   gsb_pose_grad                  <- autograd of the bmm at src/Render.cc:750-752 into Tcw
 *   gsb_adam_step                  <- torch::optim::Adam step of src/Gaussian.cc:131-175
 *
 * Conventions
 *   - All data pointers are DEVICE pointers to contiguous fp32 / int32 arrays unless the
 *     name ends in _host.  Optional inputs are NULL (the reference passes the null
 *     data_ptr of an empty tensor: include/Rasterizer.cuh:320-334).
 *   - Matrices are 16 floats read column-major (m[0]x + m[4]y + m[8]z + m[12]), exactly as
 *     the reference kernels read them (auxiliary.h:58-77).
 *   - Every call is asynchronous on `stream` (a cudaStream_t; NULL = legacy default stream)
 *     except where noted; the library keeps no global mutable state, never calls
 *     cudaMalloc/cudaFree on the hot path, and is re-entrant from multiple threads.
 *   - Functions return GSB_OK (0) or a negative gsb_status; they never throw.
 *     gsb_last_error() returns a thread-local description of the last failure.
 *   - Gradient outputs are FULLY WRITTEN by gsb_backward (zero for invisible Gaussians);
 *     the caller does not need to zero them (the reference requires zeroed buffers,
 *     src/Rasterizer.cu:253-261; zeroed buffers remain valid input).
 */
#ifndef GSB_H_INCLUDED
#define GSB_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define GSB_VERSION_MAJOR 0
#define GSB_VERSION_MINOR 2

typedef enum gsb_status {
    GSB_OK = 0,
    GSB_ERR_INVALID_ARGUMENT = -1,  /* reference: AT_ERROR / std::invalid_argument (Rasterizer.cu:158, Rasterizer.cuh:310-316) */
    GSB_ERR_CUDA = -2,              /* a CUDA runtime call or launch failed */
    GSB_ERR_WORKSPACE = -3,         /* caller-provided workspace too small / allocator returned NULL */
    GSB_ERR_OVERFLOW = -4,          /* num_rendered exceeded the binning capacity given to gsb_forward_ws */
    GSB_ERR_UNSUPPORTED = -5        /* e.g. NUM_CHANNELS != 3 without precomputed colours (impl.cu:245-248) */
} gsb_status;

typedef void* gsb_stream_t; /* cudaStream_t */

/* Scratch allocator callback: return >= bytes of device memory (256-byte aligned), or
 * NULL on failure.  Mirrors std::function<char*(size_t)> of rasterizer.h:32-34; unlike
 * the reference's resizeFunctional (src/Rasterizer.cu:127-134) the memory need NOT be
 * zeroed.  The returned block must stay alive until gsb_backward has consumed it. */
typedef void* (*gsb_alloc_fn)(void* user, size_t bytes);

/* Inputs of one rasterization (argument set of rasterizer.h:35-53). */
typedef struct gsb_raster_args {
    int P;                       /* number of Gaussians */
    int D;                       /* active SH degree (0..3) */
    int M;                       /* SH coefficients per Gaussian (0 if colours are precomputed) */
    int width, height;
    const float* background;     /* [3] */
    const float* means3D;        /* [P,3] */
    const float* shs;            /* [P,M,3] or NULL */
    const float* colors_precomp; /* [P,3] or NULL (exactly one of shs / colors_precomp) */
    const float* opacities;      /* [P] */
    const float* scales;         /* [P,3] or NULL */
    float scale_modifier;
    const float* rotations;      /* [P,4] (w,x,y,z), used as given, NOT normalised (forward.cu:127); 16-byte aligned */
    const float* cov3D_precomp;  /* [P,6] or NULL (exactly one of scales+rotations / cov3D_precomp); 8-byte aligned */
    const float* viewmatrix;     /* [16] */
    const float* projmatrix;     /* [16] */
    const float* cam_pos;        /* [3] (only read on the SH path) */
    float tan_fovx, tan_fovy;
    int prefiltered;             /* reference traps if a culled point was declared prefiltered; here: ignored */
    /* Tile-row shard (multi-GPU; not in the reference): only the 16-pixel tile rows
     * [tile_row_begin, tile_row_end) are binned, sorted and blended; pixels outside the band are NOT written,
     * gradients hold this band's share (summing the bands' images and gradients gives the full frame).
     * radii are unaffected.  0, 0 = the whole image (what a zero-initialised struct asks for); an EMPTY band is therefore
     * passed as (k, k) with k > 0, e.g. (tile rows of the image, same): nothing is binned or blended, all gradients are 0. */
    int tile_row_begin, tile_row_end;
} gsb_raster_args;

/* Gradient outputs of rasterizer.h:74-83.  Any pointer may be NULL to skip that output. */
typedef struct gsb_grad_outputs {
    float* dL_dmean2D;   /* [P,3]  (x,y written, z = 0)                               */
    float* dL_dconic;    /* [P,4]  (slots 0,1,3 written, slot 2 = 0; backward.cu:549-551) */
    float* dL_dopacity;  /* [P]    */
    float* dL_dcolor;    /* [P,3]  */
    float* dL_dmean3D;   /* [P,3]  */
    float* dL_dcov3D;    /* [P,6]  */
    float* dL_dsh;       /* [P,M,3] (only when shs != NULL) */
    float* dL_dscale;    /* [P,3]  */
    float* dL_drot;      /* [P,4]  w.r.t. the quaternion as given (no normalisation Jacobian; backward.cu:340) */
} gsb_grad_outputs;

/* ---- library info ------------------------------------------------------------------- */
int gsb_version(void);                  /* (major << 16) | minor */
const char* gsb_last_error(void);       /* thread-local; "" if none */

/* ---- workspace sizing ----------------------------------------------------------------
 * Byte sizes of the three opaque state blobs (the analogue of required<GeometryState>(P),
 * required<ImageState>(W*H), required<BinningState>(R), rasterizer_impl.h:67-73). */
size_t gsb_geometry_bytes(int P);
size_t gsb_image_bytes(int width, int height);
size_t gsb_binning_bytes(long long max_rendered);
/* Same three numbers through one call (out pointers may be NULL). */
int gsb_workspace_query(int P, int width, int height, long long max_rendered,
                        size_t* geometry_bytes, size_t* image_bytes, size_t* binning_bytes);

/* ---- forward ------------------------------------------------------------------------
 * Drop-in for Rasterizer::forward.  Allocates the three state blobs through the
 * callbacks, performs ONE stream synchronisation to learn num_rendered (the reference
 * does a blocking cudaMemcpy at rasterizer_impl.cu:285) and returns it (>= 0), or a
 * negative gsb_status.  out_color [3,H,W], out_depth [1,H,W], radii [P] (may be NULL)
 * are fully written. */
int gsb_forward(const gsb_raster_args* args,
                gsb_alloc_fn geometry_alloc, void* geometry_user,
                gsb_alloc_fn binning_alloc, void* binning_user,
                gsb_alloc_fn image_alloc, void* image_user,
                float* out_color, float* out_depth, int* radii,
                gsb_stream_t stream);

/* Sync-free forward over caller-provided workspaces (sizes from gsb_*_bytes).  The
 * binning blob is sized for `max_rendered` tile instances; if the frame needs more, the
 * instance list is truncated on the device, an overflow flag is latched in the geometry
 * blob and gsb_num_rendered() reports GSB_ERR_OVERFLOW.  Returns GSB_OK once everything
 * is enqueued. */
int gsb_forward_ws(const gsb_raster_args* args,
                   void* geometry, size_t geometry_bytes,
                   void* binning, size_t binning_bytes, long long max_rendered,
                   void* image, size_t image_bytes,
                   float* out_color, float* out_depth, int* radii,
                   gsb_stream_t stream);

/* Synchronises `stream` and returns the num_rendered recorded in a geometry blob
 * (>= 0), or GSB_ERR_OVERFLOW. */
long long gsb_num_rendered(const void* geometry, gsb_stream_t stream);

/* ---- backward -----------------------------------------------------------------------
 * Drop-in for Rasterizer::backward.  `R` is the value returned by gsb_forward (pass -1
 * after gsb_forward_ws: the count is read from the geometry blob on the device).
 * `radii` may be NULL (the blob keeps a copy).  dL_dpix is [3,H,W]; the gradient of the
 * depth output is not propagated (include/Rasterizer.cuh:210). */
int gsb_backward(const gsb_raster_args* args, long long R, const int* radii,
                 const void* geometry, const void* binning, const void* image,
                 const float* dL_dpix, const gsb_grad_outputs* grads,
                 gsb_stream_t stream);

/* ---- fused RGB + depth / silhouette pass (extension) ---------------------
 * Every optimisation iteration of the reference rasterizes the SAME geometry twice: once with the
 * Gaussians' colours and once with colours [z_cam, 1, 0] (depth and silhouette; src/Render.cc:445-448,
 * :949-981), duplicating projection, binning, sort and every alpha evaluation.  These entry points
 * blend five channels in one pass:
 *   out_color        [3,H,W]  = the RGB pass' colour output;
 *   out_depth_sil    [2,H,W]  = channels 0 / 1 of the depth pass (sum z alpha T + T bg[0], sum alpha T + T bg[1]),
 *                               with z the view-space depth of the Gaussian (= z_cam in the reference's default
 *                               mode: identity view matrix, pre-transformed means);
 *   out_median_depth [1,H,W]  = the third output of either pass ("renderedSurdepth");
 * bit-identical to the two separate passes.  The backward takes dL/d(out_color) and dL/d(out_depth_sil) and
 * returns the SUM of the two passes' gradients (what autograd accumulates), plus dL_dzcolor [P] (may be NULL):
 * the gradient of the z_cam colour.  With z_attached != 0 that gradient is also added to dL_dmean3D[:, 2] -- the
 * colour is a function of the means in mapping mode (identity view matrix: z_cam = mean.z); 0 reproduces the
 * detached colour of tracking mode (src/Render.cc:957). */
int gsb_forward_fused_ws(const gsb_raster_args* args,
                         void* geometry, size_t geometry_bytes,
                         void* binning, size_t binning_bytes, long long max_rendered,
                         void* image, size_t image_bytes,
                         float* out_color, float* out_depth_sil, float* out_median_depth, int* radii,
                         gsb_stream_t stream);
int gsb_backward_fused(const gsb_raster_args* args, const int* radii,
                       const void* geometry, const void* binning, const void* image,
                       const float* dL_dcolor, const float* dL_ddepth_sil,
                       const gsb_grad_outputs* grads, float* dL_dzcolor, int z_attached, gsb_stream_t stream);

/* ---- the per-Gaussian tail of a mapping iteration in ONE launch (extension) ---------------------------
 * What follows the rasterizer's per-pixel backward in Render::RenderForFrame (src/Render.cc:420-476), per Gaussian:
 * BACKWARD::preprocess (backward.cu:560-621) -> autograd of the activation prologue (src/Render.cc:750-759: Tcw [mean;1],
 * sigmoid, normalize, exp) -> the scale regularisers (:462-469) -> torch::optim::Adam with one learning rate per tensor
 * (src/Gaussian.cc:131-175).  gsb_backward_fused + gsb_prologue_backward + gsb_scale_regulariser + gsb_adam_step_groups do the
 * same in four streaming passes; here one thread carries a Gaussian from the blend-backward sums to its updated parameters and
 * no gradient array is written (unless `grads` asks for the raw-parameter gradients as well).
 * Group order everywhere: 0 means [P,3] (world frame), 1 rgb [P,3], 2 logit opacities [P], 3 log scales [P,3],
 * 4 unnormalised quaternions [P,4].  The forward must have been run on the prologue's outputs of THESE parameters with
 * colors_precomp = params[1], no SH, no precomputed covariances, over the whole image.
 * If that forward overflowed its binning blob (the lists it blended were truncated) NOTHING is updated: the caller grows the
 * blob, renders again and calls this again with the same `step`. */
typedef struct gsb_map_update {
    const float* Tcw;          /* [4,4] row-major, device: the pose the prologue used */
    float* params[5];          /* updated in place */
    float* exp_avg[5];         /* Adam first moments, updated in place */
    float* exp_avg_sq[5];      /* Adam second moments, updated in place */
    float* grads[5];           /* all NULL, or all set: dL/d(raw parameter) is written too (regularisers included) */
    float lr[5];
    double beta1, beta2, eps;
    long long step;            /* 1-based count of THIS Adam step (bias corrections) */
    float* dL_dTcw;            /* [3,4] device or NULL: sum_i g_i [mean_i; 1]^T */
    float max_scalar;          /* scale regularisers as in gsb_scale_regulariser; <= 0: off */
    float w_scalar, w_long;
    float* reg_terms;          /* 8 floats, device (required when the regularisers are on): as gsb_scale_regulariser's `terms` */
} gsb_map_update;
/* dL_ddepth_sil NULL: the three-channel pass (gsb_forward_ws); otherwise the five-channel pass (gsb_forward_fused_ws). */
int gsb_backward_fused_update(const gsb_raster_args* args, const int* radii,
                              const void* geometry, const void* binning, const void* image,
                              const float* dL_dcolor, const float* dL_ddepth_sil, int z_attached,
                              const gsb_map_update* update, gsb_stream_t stream);

/* Tracking (Render::RenderStartTraking, src/Render.cc:1052-1127) optimises the camera pose only: the per-pixel backward followed by
 * ONE per-Gaussian kernel that goes as far as dL/dmean_cam and reduces dL_dTcw [3,4] = sum_i dL/dmean_cam_i [means_world_i; 1]^T
 * on chip -- what gsb_backward_fused + gsb_prologue_backward deliver, without writing any per-Gaussian gradient.
 * args as for the forward (means3D = the camera-frame means the prologue wrote); a tile-row band yields that band's partial sum.
 * dL_ddepth_sil NULL: three-channel pass. */
int gsb_backward_fused_pose(const gsb_raster_args* args, const int* radii,
                            const void* geometry, const void* binning, const void* image,
                            const float* dL_dcolor, const float* dL_ddepth_sil, int z_attached,
                            const float* means_world, float* dL_dTcw, gsb_stream_t stream);

/* ---- visibility helpers --------------------------------------------------------------*/
/* Radii-only projection (Rasterizer::visible_filter): radii[P] fully written. */
int gsb_visible_filter(const gsb_raster_args* args, int* radii, gsb_stream_t stream);
/* present[i] = (view-space z > 0.2) (Rasterizer::markVisible / checkFrustum). */
int gsb_mark_visible(int P, const float* means3D, const float* viewmatrix,
                     const float* projmatrix, uint8_t* present, gsb_stream_t stream);

/* ---- simple_knn -----------------------------------------------------------------------
 * mean_dist2[i] = mean of the squared distances from points[i] to its 3 nearest
 * neighbours (SimpleKNN::knn).  `workspace` must hold gsb_knn_workspace_bytes(P) bytes.
 * Performs one stream synchronisation (the reference does two blocking copies of the
 * bounding box, simple_knn.cu:193-200). */
size_t gsb_knn_workspace_bytes(int P);
int gsb_knn_mean_dist2(int P, const float* points, float* mean_dist2,
                       void* workspace, size_t workspace_bytes, gsb_stream_t stream);

/* ---- fused caller-side pieces (extensions) ------------------------------*/
/* Prologue of Render::StartSplatting (src/Render.cc:750-759) in one pass:
 *   means_cam = (Tcw * [mean;1]).xyz, opacities = sigmoid(logit), rot = normalize(q),
 *   scales = exp(log_scale).  Tcw is [4,4] ROW-major (torch layout).  Any output may be NULL. */
int gsb_prologue(int P, const float* Tcw, const float* means_world, const float* logit_opacities,
                 const float* unnorm_quats, const float* log_scales,
                 float* means_cam, float* opacities, float* rotations, float* scales,
                 gsb_stream_t stream);
/* Backward of that prologue: chain rule through sigmoid / normalize / exp / the rigid
 * transform, plus dL/dTcw[0:3,:] = sum_i g_i [p_i;1]^T (the "camera-pose backward";
 * src/Render.cc:750-752 differentiated by autograd in the reference).  dL_dTcw is 12 floats
 * (rows 0..2 of the 4x4, row-major), fully written.  Any output may be NULL. */
int gsb_prologue_backward(int P, const float* Tcw, const float* means_world,
                          const float* logit_opacities, const float* unnorm_quats,
                          const float* log_scales,
                          const float* dL_dmeans_cam, const float* dL_dopacities,
                          const float* dL_drotations, const float* dL_dscales,
                          float* dL_dmeans_world, float* dL_dlogit_opacities,
                          float* dL_dunnorm_quats, float* dL_dlog_scales, float* dL_dTcw,
                          gsb_stream_t stream);
/* dL/dTcw only (12 floats), from camera-frame mean gradients. */
int gsb_pose_grad(int P, const float* means_world, const float* dL_dmeans_cam, float* dL_dTcw,
                  gsb_stream_t stream);

/* Fused multi-tensor Adam (torch::optim::Adam semantics: bias-corrected, eps outside the
 * sqrt, no weight decay / amsgrad; src/Gaussian.cc:131-175).  One launch updates n
 * contiguous fp32 values: p -= lr * mhat / (sqrt(vhat) + eps).  `step` is the 1-based step
 * count AFTER this update.  The hyper-parameters are doubles, as torch holds them: 1 - beta and lr / bias_correction are
 * formed in double and rounded once (1.0f - 0.999f is off by 1.3e-5 relative). */
int gsb_adam_step(long long n, float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                  double lr, double beta1, double beta2, double eps, long long step,
                  gsb_stream_t stream);

/* ---- fused loss of one mapping iteration (extension) ------------------------
 * The producer of dL/dpixel in Render::RenderForFrame (src/Render.cc:454-469) and its autograd, in two kernels:
 *   loss = w_image * (lambda * mean|I - G| + (1 - lambda) * (1 - SSIM(I, G)))
 *        + w_depth * mean_{G_d > 0} |D - G_d|  +  w_surdepth * mean_{G_d > 0, S > 0.99} |M - G_d|
 * with I = color [3,H,W], (D, S) = depth_sil [2,H,W], M = median_depth [1,H,W] (no gradient: Rasterizer.cuh:210),
 * G / G_d the ground-truth image / depth; SSIM exactly as src/Utils.cc:68-100 builds it (11x11 window from the
 * reference's off-centre 1-D weights, zero padding, C1 = 0.01^2, C2 = 0.03^2, global mean).  depth_sil, median_depth,
 * gt_depth, dL_ddepth_sil may be NULL (image term only).  Outputs: dL_dcolor [3,H,W], dL_ddepth_sil [2,H,W] and
 * loss_terms[8] on the DEVICE = {l1, ssim, depth_l1, surdepth_l1, total, n_valid, n_valid_sur, 0}.  The scale
 * regularisers of Render.cc:462-467 act on the parameters, not on pixels, and stay with the caller. */
size_t gsb_loss_scratch_bytes(int width, int height);
int gsb_mapping_loss(int width, int height, const float* color, const float* depth_sil, const float* median_depth,
                     const float* gt_color, const float* gt_depth,
                     float lambda_, float w_image, float w_depth, float w_surdepth,
                     float* dL_dcolor, float* dL_ddepth_sil, float* loss_terms,
                     void* scratch, size_t scratch_bytes, gsb_stream_t stream);

/* ---- fused loss of one tracking iteration (extension) -------------------------
 * The producer of dL/dpixel in Render::RenderStartTraking (src/Render.cc:1075-1093; L1LossForTracking, src/Utils.cc:45-52) and its
 * autograd, in one kernel:  mask = silhouette > 0.99 and gt_depth is not NaN ("uncertainDepth");
 *   loss = w_image * sum_mask |I - G| (three channels) + w_depth * sum_mask |D - G_d|,
 * D = median_depth [H,W] when use_surdepth (it carries no gradient: include/Rasterizer.cuh:210), else depth_sil[0].
 * Outputs: dL_dcolor [3,H,W] = w_image sign(I - G) mask, dL_ddepth_sil [2,H,W] (may be NULL; channel 0 = w_depth sign(D - G_d) mask
 * without use_surdepth, zero otherwise) and loss_terms[8] on the DEVICE = {image_l1, depth_l1, loss, n_mask, 0, 0, 0, internal}.
 * The ORB reprojection term of the same loop acts on the pose, not on pixels, and stays with the caller. */
int gsb_tracking_loss(int width, int height, const float* color, const float* depth_sil, const float* median_depth,
                      const float* gt_color, const float* gt_depth, float w_image, float w_depth, int use_surdepth,
                      float* dL_dcolor, float* dL_ddepth_sil, float* loss_terms, gsb_stream_t stream);

/* The two scale regularisers of the mapping loss (src/Render.cc:462-469), which act on the parameters, not on pixels:
 *   big = where(exp(log_scales) > max_scalar)[0]   (a row appears once per axis that exceeds),
 *   reg_scalar = sum_big (max_axis exp(ls) - max_scalar),  reg_long = mean_big (max_axis exp(ls) - min_axis exp(ls)),
 * max_scalar = 0.1 * scene radius (:418).  ADDS d(w_scalar reg_scalar + w_long reg_long)/d(log_scales) to dL_dlog_scales
 * [P,3] (may be NULL) and writes terms[0..3] = {reg_scalar, reg_long (NaN when nothing is selected, as in the reference),
 * number of selected (row, axis) pairs, 0}; `terms` is 8 floats of DEVICE memory (4..7 are scratch). */
int gsb_scale_regulariser(int P, const float* log_scales, float max_scalar, float w_scalar, float w_long,
                          float* dL_dlog_scales, float* terms, gsb_stream_t stream);

/* ---- densification: back-projection of selected pixels (extension) ------
 * GPU twin of the host loops Render::ProjectPixel / Render::InitGaussianPoint (src/Render.cc:617-655, :666-707) and of the
 * parameter initialisation of Gaussian::AddGaussianPoints (src/Gaussian.cc:50-74, SinglePixel scales).  For every pixel
 * with mask >= 250 (mask NULL: every pixel) and depth > 0, in row-major pixel order (= the ids the reference assigns):
 *   mean = Twc [((j-cx) z/fx, (i-cy) z/fy, z); 1], rgb = image[:, i, j], log_scale = log(|mean.z| / ((fx+fy)/2)) x 3,
 *   unnorm_quat = (1,0,0,0), logit_opacity = 1.
 * mask [H,W] u8, depth [H,W], image [3,H,W] are DEVICE pointers; Twc_host is 16 floats, row-major, on the HOST.  Rows
 * beyond `capacity` are dropped; *count (DEVICE int) receives the number of selected pixels either way.  max_z
 * (DEVICE float, may be NULL) is raised to the largest selected depth (Render::mMaxZ; initialise it to the running
 * value).  Any of rgb / log_scales / unnorm_quats / logit_opacities may be NULL. */
size_t gsb_backproject_scratch_bytes(int width, int height);
int gsb_backproject(int width, int height, const uint8_t* mask, const float* depth, const float* image,
                    float fx, float fy, float cx, float cy, const float* Twc_host, int capacity,
                    float* means, float* rgb, float* log_scales, float* unnorm_quats, float* logit_opacities,
                    int* count, float* max_z, void* scratch, size_t scratch_bytes, gsb_stream_t stream);

/* ---- prune (extension) ---------------------------------------------------
 * gsb_low_opacity_keep: keep[i] = !(sigmoid(logit_opacity[i]) < threshold)  (Gaussian::RemoveLowOpcitiesGaussian, pruneOpcities 0.005).
 * gsb_prune_rows: order-preserving compaction of up to 16 row-major per-Gaussian tensors [P, width_k] -> [K, width_k] by the
 * keep flags in ONE pass -- what Gaussian::RemovePoints / PruneOptimizer (src/Gaussian.cc:209-239) do with one index_select per
 * parameter tensor and per Adam moment.  src_host / dst_host / widths_host are HOST tables of DEVICE pointers / widths (dst != src);
 * *count (DEVICE int) receives K. */
size_t gsb_prune_scratch_bytes(int P);
int gsb_low_opacity_keep(int P, const float* logit_opacities, float threshold, uint8_t* keep, gsb_stream_t stream);
int gsb_prune_rows(int P, const uint8_t* keep, int ntensors, const float* const* src_host, float* const* dst_host,
                   const int* widths_host, int* count, void* scratch, size_t scratch_bytes, gsb_stream_t stream);

/* ---- multi-GPU exchange step (the reference is single-GPU) ----------------------
 * In-place SUM all-reduce of n fp32 values (n % 4 == 0) that live at the same offset of a
 * symmetric, peer-mapped allocation on every rank of one NVLink / NVSwitch box -- the packed
 * [14, P] gradient block between loss.backward() and the Adam step (src/Render.cc:471-475).
 * ONE kernel per rank: pairwise cross-rank barrier, reduce this rank's 1/world slice
 * (multimem.ld_reduce through the NVSwitch multicast mapping when multicast_ptr != NULL, else
 * 128-bit peer loads in rank order), publish it to every rank (multimem.st / peer stores), system
 * fence, barrier.  All ranks end up bit-identical.  Every rank must call it in the same order.
 *   peer_ptrs: HOST array of `world` DEVICE pointers -- this rank's mapping of every rank's buffer
 *              (its own at index `rank`), each 16-byte aligned;
 *   sync_ptrs: HOST array of `world` DEVICE pointers to a symmetric scratch of
 *              gsb_exchange_sync_bytes(world) bytes per rank, zeroed ONCE when it is allocated.
 * world == 1 is a no-op.  The mappings come from the caller (CUDA VMM / IPC; the Python host side
 * uses torch.distributed._symmetric_memory). */
size_t gsb_exchange_sync_bytes(int world);
int gsb_exchange_allreduce(void* multicast_ptr, void* const* peer_ptrs, void* const* sync_ptrs,
                           long long n, int rank, int world, gsb_stream_t stream);

/* The same update over a packed block of `ngroups` (<= 8) contiguous parameter groups with one learning rate each
 * (one torch param group per tensor, src/Gaussian.cc:158-175) in ONE launch.  group_sizes_host / lrs_host are HOST arrays;
 * param / grad / exp_avg / exp_avg_sq hold sum(group_sizes) values. */
int gsb_adam_step_groups(int ngroups, const long long* group_sizes_host, const float* lrs_host,
                         float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                         double beta1, double beta2, double eps, long long step, gsb_stream_t stream);

/* ---- host-buffer convenience (bench "e2e" leg and quick integration tests) -------------
 * One forward + backward with every array in HOST memory (pinned recommended): copies the
 * inputs H2D, runs gsb_forward_ws + gsb_backward, copies colour/depth/radii and the
 * gradients back.  `device_scratch` must hold gsb_host_scratch_bytes(...) bytes of DEVICE
 * memory.  Synchronises `stream` before returning.  Returns num_rendered or an error. */
size_t gsb_host_scratch_bytes(int P, int M, int width, int height, long long max_rendered);
long long gsb_forward_backward_host(const gsb_raster_args* host_args, long long max_rendered,
                                    const float* dL_dpix_host,
                                    float* out_color_host, float* out_depth_host, int* radii_host,
                                    const gsb_grad_outputs* host_grads,
                                    void* device_scratch, size_t device_scratch_bytes,
                                    gsb_stream_t stream);
/* The same frame, fully ASYNCHRONOUS: kernels on `stream`, uploads / downloads on the library's per-thread upload / download
 * streams (chained to `stream` by events; `stream` completes only after the frame's last download); returns at once.
 * `status_host` (3 x uint32 of PINNED host memory) receives {num_rendered, instances actually binned, overflow latch} when
 * `stream` gets there (overflow != 0: the frame was truncated to max_rendered instances -- grow the scratch and resubmit).
 * Meant for K frames in flight -- K device scratch sets, K streams, K sets of host output buffers: PCIe is full duplex and
 * the copy engines are separate from the SMs, so frame i's gradient download runs under frame i+1's upload and frame i+2's
 * kernels (gsorb_slam_b200/host.py: HostPipeline).  The host arrays must stay valid until `stream` has drained. */
int gsb_forward_backward_host_async(const gsb_raster_args* host_args, long long max_rendered,
                                    const float* dL_dpix_host,
                                    float* out_color_host, float* out_depth_host, int* radii_host,
                                    const gsb_grad_outputs* host_grads,
                                    void* device_scratch, size_t device_scratch_bytes,
                                    unsigned int* status_host, gsb_stream_t stream);

/* ---- introspection for parity tests ----------------------------------------------------
 * Copies of internal state into caller DEVICE buffers (any pointer may be NULL):
 * final transmittance [H*W], n_contrib [H*W], tile ranges [tiles*2], sorted instance list
 * [R], projected state per Gaussian (depths [P], means2D [P,2], conic_opacity [P,4],
 * tiles_touched [P]). */
int gsb_debug_image_state(const void* image, int width, int height,
                          float* final_T, uint32_t* n_contrib, uint32_t* ranges,
                          gsb_stream_t stream);
/* Number of (pixel, splat) pairs the last forward pass blended = set bits of the hit words it recorded (one word per
 * 32-entry window of a tile's list and pixel), summed over every pixel's windows up to its n_contrib; `count` is one
 * DEVICE uint64.  The measurement unit of the blend kernels (bench.py: warp-instructions per blended pair). */
int gsb_debug_blended_pairs(const void* geometry, const void* binning, const void* image, int width, int height,
                            unsigned long long* count, gsb_stream_t stream);
int gsb_debug_binning_state(const void* geometry, const void* binning, long long R,
                            uint32_t* point_list, gsb_stream_t stream);
int gsb_debug_geometry_state(const void* geometry, int P, float* depths, float* means2D,
                             float* conic_opacity, uint32_t* tiles_touched,
                             gsb_stream_t stream);

/* Number of kernels launched by this thread through the library since the last call
 * (used by bench.py for its gpu_launches claim). */
long long gsb_launch_count_reset(void);

/* Per-stage device timing for bench.py's roofline line.  Between gsb_profile_begin() and
 * gsb_profile_end() every kernel launched by this thread through the library is bracketed by
 * CUDA events on its own stream.  gsb_profile_end synchronises those events and returns, per
 * stage, the summed milliseconds and the number of timed launches (arrays of gsb_num_stages()
 * entries, either may be NULL). */
int gsb_num_stages(void);
const char* gsb_stage_name(int stage);
int gsb_profile_begin(void);
int gsb_profile_end(float* stage_ms, int* stage_count);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* GSB_H_INCLUDED */
