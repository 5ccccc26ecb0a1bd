// api.cu -- the C ABI of libgsb.so (include/gsb.h): argument validation, workspace layout,
// kernel orchestration.  No torch types, no global mutable state (error text and launch
// counter are thread-local), no cudaMalloc/cudaFree on any path.
//
// Orchestration replaces CudaRasterizer::Rasterizer::{forward,backward,visible_filter,
// markVisible} (rasterizer_impl.cu:199-345, :405-498, :348-401, :142-154).
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>
#include "common.cuh"
#include "stage.cuh"

namespace gsb {

static thread_local char g_err[512] = "";
static thread_local long long g_launches = 0;

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches += n; }

struct ProfEvent { int stage; cudaEvent_t a, b; };
static thread_local bool g_prof_on = false;
static thread_local std::vector<ProfEvent>* g_prof = nullptr;

StageTimer::StageTimer(int st, cudaStream_t str) : stage(st), s(str), ev(nullptr)
{
    if (!g_prof_on) return;
    ProfEvent e;
    e.stage = st;
    if (cudaEventCreate(&e.a) != cudaSuccess || cudaEventCreate(&e.b) != cudaSuccess) return;
    cudaEventRecord(e.a, s);
    g_prof->push_back(e);
    ev = reinterpret_cast<void*>(g_prof->size());  // index + 1
}
StageTimer::~StageTimer()
{
    if (!ev) return;
    cudaEventRecord((*g_prof)[reinterpret_cast<size_t>(ev) - 1].b, s);
}

static int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// Mirrors the argument checks of the reference surface: means3D shape (src/Rasterizer.cu:158-160),
// sh / colour and scale+rotation / cov3D exclusivity (include/Rasterizer.cuh:310-316), NUM_CHANNELS
// (rasterizer_impl.cu:245-248).
static int validate(const gsb_raster_args* a, bool need_colors, bool need_opacity = true)
{
    if (!a) return fail(GSB_ERR_INVALID_ARGUMENT, "args is NULL");
    if (a->P < 0) return fail(GSB_ERR_INVALID_ARGUMENT, "P must be >= 0 (got %d)", a->P);
    if (a->width <= 0 || a->height <= 0)
        return fail(GSB_ERR_INVALID_ARGUMENT, "image size must be positive (got %dx%d)", a->width, a->height);
    if (!(a->tan_fovx > 0.f) || !(a->tan_fovy > 0.f)) return fail(GSB_ERR_INVALID_ARGUMENT, "tan_fov must be positive");
    if (a->tile_row_begin < 0 || a->tile_row_end < a->tile_row_begin || a->tile_row_end > (a->height + TILE_Y - 1) / TILE_Y)
        return fail(GSB_ERR_INVALID_ARGUMENT, "tile row band [%d, %d) is outside the image's %d tile rows", a->tile_row_begin,
                    a->tile_row_end, (a->height + TILE_Y - 1) / TILE_Y);
    if (!a->viewmatrix || !a->projmatrix) return fail(GSB_ERR_INVALID_ARGUMENT, "viewmatrix / projmatrix are required");
    if (a->P > 0) {
        if (!a->means3D) return fail(GSB_ERR_INVALID_ARGUMENT, "means3D must have dimensions (num_points, 3)");
        const bool has_sr = a->scales && a->rotations;
        if ((a->scales != nullptr) != (a->rotations != nullptr))
            return fail(GSB_ERR_INVALID_ARGUMENT, "scales and rotations must be given together");
        if (has_sr == (a->cov3D_precomp != nullptr))
            return fail(GSB_ERR_INVALID_ARGUMENT, "Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!");
        // the kernels read rotations with 128-bit and cov3D_precomp with 64-bit loads (preprocess.cu, gauss_bwd.cu): a
        // misaligned pointer would be a sticky CUDA fault for the whole context, so it is refused here (include/gsb.h)
        if (a->rotations && (reinterpret_cast<uintptr_t>(a->rotations) & 15))
            return fail(GSB_ERR_INVALID_ARGUMENT, "rotations must be 16-byte aligned");
        if (a->cov3D_precomp && (reinterpret_cast<uintptr_t>(a->cov3D_precomp) & 7))
            return fail(GSB_ERR_INVALID_ARGUMENT, "cov3D_precomp must be 8-byte aligned");
        if (need_colors) {
            if (need_opacity && !a->opacities) return fail(GSB_ERR_INVALID_ARGUMENT, "opacities are required");
            if ((a->shs != nullptr) == (a->colors_precomp != nullptr))
                return fail(GSB_ERR_INVALID_ARGUMENT, "Please provide exactly one of either SHs or precomputed colors!");
            if (a->shs) {
                if (a->D < 0 || a->D > 3) return fail(GSB_ERR_INVALID_ARGUMENT, "SH degree must be 0..3 (got %d)", a->D);
                if (a->M < (a->D + 1) * (a->D + 1))
                    return fail(GSB_ERR_INVALID_ARGUMENT, "M = %d SH coefficients cannot hold degree %d", a->M, a->D);
                if (!a->cam_pos) return fail(GSB_ERR_INVALID_ARGUMENT, "cam_pos is required with SHs");
            }
        }
    }
    if (need_colors && !a->background) return fail(GSB_ERR_INVALID_ARGUMENT, "background is required");
    return GSB_OK;
}

// host-buffer entry points: the arrays are copied into 256-byte aligned device scratch, so host alignment is irrelevant
static int validate_host(const gsb_raster_args* a)
{
    if (!a) return fail(GSB_ERR_INVALID_ARGUMENT, "args is NULL");
    gsb_raster_args t = *a;
    if (t.rotations) t.rotations = reinterpret_cast<const float*>(uintptr_t(256));
    if (t.cov3D_precomp) t.cov3D_precomp = reinterpret_cast<const float*>(uintptr_t(256));
    return validate(&t, true);
}

static FwdParams make_params(const gsb_raster_args* a)
{
    FwdParams p;
    p.P = a->P; p.D = a->D; p.M = a->shs ? a->M : 0; p.W = a->width; p.H = a->height;
    p.tiles_x = (a->width + TILE_X - 1) / TILE_X;
    p.tiles_y = (a->height + TILE_Y - 1) / TILE_Y;
    const bool whole = a->tile_row_begin == 0 && a->tile_row_end == 0;
    p.band_y0 = whole ? 0 : a->tile_row_begin;
    p.band_y1 = whole ? p.tiles_y : a->tile_row_end;
    p.background = a->background; p.means3D = a->means3D; p.shs = a->shs; p.colors_precomp = a->colors_precomp;
    p.opacities = a->opacities; p.scales = a->scales; p.scale_modifier = a->scale_modifier; p.rotations = a->rotations;
    p.cov3D_precomp = a->cov3D_precomp; p.viewmatrix = a->viewmatrix; p.projmatrix = a->projmatrix; p.cam_pos = a->cam_pos;
    p.tan_fovx = a->tan_fovx; p.tan_fovy = a->tan_fovy;
    p.focal_y = a->height / (2.0f * a->tan_fovy);  // rasterizer_impl.cu:224-225
    p.focal_x = a->width / (2.0f * a->tan_fovx);
    return p;
}

static int forward_stage1(const FwdParams& p, char* geom, const GeomLayout& GL, char* image, const ImageLayout& IL, int* radii,
                          uint32_t capacity, cudaStream_t s)
{
    // the header needs no clearing: the scan kernel writes every field the rasterizer reads
    GSB_CUDA_CHECK(cudaMemsetAsync(image + IL.tile_count, 0, ((size_t)IL.tiles_x * IL.tiles_y + 1) * 4 * TILE_CTR_STRIDE, s));
    if (p.P > 0) return launch_preprocess(p, geom, GL, image, IL, radii, capacity, s);   // its last CTA also scans the tile counts
    return launch_tile_scan(geom, GL, image, IL, capacity, p.P, s);                       // empty map: all-empty segments
}

static int forward_stage2(const FwdParams& p, char* geom, const GeomLayout& GL, char* binning, const BinningLayout& BL,
                          char* image, const ImageLayout& IL, long long grid_instances, float* out_color, float* out_depth,
                          float* out_depth_sil, cudaStream_t s)
{
    if (int rc = launch_binning(p, geom, GL, binning, BL, image, IL, grid_instances, s)) return rc;
    return launch_blend_forward(p, geom, GL, binning, BL, image, IL, out_color, out_depth, out_depth_sil, s);
}

// one CTA per tile, one thread per pixel in the blend kernels' pixel order (warp = 8x4 region, lane = row-major pixel of the region):
// set bits of the pixel's hit words over the windows up to its last contributor
__global__ void __launch_bounds__(BLEND_THREADS)
count_blended_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ n_contrib, const uint32_t* __restrict__ hits_tail,
                     const char* __restrict__ binning, const GeomHeader* __restrict__ hdr, int W, int H, unsigned long long* __restrict__ count)
{
    const uint32_t tile = blockIdx.y * gridDim.x + blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int px = blockIdx.x * TILE_X + (warp & 1) * 8 + (lane & 7), py = blockIdx.y * TILE_Y + (warp >> 1) * 4 + (lane >> 3);
    const uint2 range = ranges[tile];
    const BinningLayout BL = BinningLayout::make((long long)hdr->layout_capacity);
    uint32_t* hits_full = reinterpret_cast<uint32_t*>(const_cast<char*>(binning) + BL.hits);
    unsigned long long c = 0;
    if (px < W && py < H) {
        const uint32_t last = n_contrib[(size_t)py * W + px];
        for (uint32_t w = 0; w * 32 < last; w++)
            c += __popc(hit_words(hits_full, const_cast<uint32_t*>(hits_tail), tile, range.x, range.y - range.x, w)[tid]);
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0 && c) atomicAdd(count, c);
}

__global__ void unpack_geometry_kernel(int P, const SplatRec* __restrict__ rec, const int* __restrict__ radii,
                                       const uint32_t* __restrict__ tiles_touched, float* depths, float* means2D,
                                       float* conic_opacity, uint32_t* tt_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const bool vis = radii[i] > 0;
    const SplatRec r = rec[i];
    if (depths) depths[i] = vis ? r.c.w : 0.f;
    if (means2D) { means2D[2 * i] = vis ? r.a.x : 0.f; means2D[2 * i + 1] = vis ? r.a.y : 0.f; }
    if (conic_opacity) {
        conic_opacity[4 * i] = vis ? r.b.x : 0.f; conic_opacity[4 * i + 1] = vis ? r.b.y : 0.f;
        conic_opacity[4 * i + 2] = vis ? r.b.z : 0.f; conic_opacity[4 * i + 3] = vis ? r.b.w : 0.f;
    }
    if (tt_out) tt_out[i] = tiles_touched[i];
}

}  // namespace gsb

using namespace gsb;

extern "C" {

int gsb_version(void) { return (GSB_VERSION_MAJOR << 16) | GSB_VERSION_MINOR; }
const char* gsb_last_error(void) { return g_err; }
long long gsb_launch_count_reset(void)
{
    const long long n = g_launches;
    g_launches = 0;
    return n;
}

static const char* kStageNames[ST_COUNT] = {"memset", "preprocess", "scan", "duplicate", "sort_histogram", "sort_passes",
                                            "tile_sort", "blend_forward", "blend_backward", "gauss_backward", "other"};
int gsb_num_stages(void) { return ST_COUNT; }
const char* gsb_stage_name(int stage) { return stage >= 0 && stage < ST_COUNT ? kStageNames[stage] : ""; }
int gsb_profile_begin(void)
{
    if (!g_prof) g_prof = new std::vector<ProfEvent>();
    for (auto& e : *g_prof) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    g_prof->clear();
    g_prof_on = true;
    return GSB_OK;
}
int gsb_profile_end(float* stage_ms, int* stage_count)
{
    g_prof_on = false;
    if (stage_ms) for (int i = 0; i < ST_COUNT; i++) stage_ms[i] = 0.f;
    if (stage_count) for (int i = 0; i < ST_COUNT; i++) stage_count[i] = 0;
    if (!g_prof) return GSB_OK;
    int rc = GSB_OK;
    for (auto& e : *g_prof) {
        float ms = 0.f;
        if (cudaEventSynchronize(e.b) != cudaSuccess || cudaEventElapsedTime(&ms, e.a, e.b) != cudaSuccess) {
            rc = fail(GSB_ERR_CUDA, "profile_end: event query failed");
        } else {
            if (stage_ms) stage_ms[e.stage] += ms;
            if (stage_count) stage_count[e.stage] += 1;
        }
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    g_prof->clear();
    return rc;
}

size_t gsb_geometry_bytes(int P) { return GeomLayout::make(P).total; }
size_t gsb_image_bytes(int width, int height) { return ImageLayout::make(width > 0 ? width : 1, height > 0 ? height : 1).total; }
size_t gsb_binning_bytes(long long max_rendered) { return BinningLayout::make(max_rendered).total; }
int gsb_workspace_query(int P, int width, int height, long long max_rendered, size_t* geometry_bytes, size_t* image_bytes,
                        size_t* binning_bytes)
{
    if (P < 0 || width <= 0 || height <= 0 || max_rendered < 0)
        return fail(GSB_ERR_INVALID_ARGUMENT, "workspace_query: bad sizes");
    if (geometry_bytes) *geometry_bytes = gsb_geometry_bytes(P);
    if (image_bytes) *image_bytes = gsb_image_bytes(width, height);
    if (binning_bytes) *binning_bytes = gsb_binning_bytes(max_rendered);
    return GSB_OK;
}

int gsb_forward(const gsb_raster_args* args, gsb_alloc_fn geometry_alloc, void* geometry_user, gsb_alloc_fn binning_alloc,
                void* binning_user, gsb_alloc_fn image_alloc, void* image_user, float* out_color, float* out_depth,
                int* radii, gsb_stream_t stream)
{
    if (int rc = validate(args, true)) return rc;
    if (!geometry_alloc || !binning_alloc || !image_alloc) return fail(GSB_ERR_INVALID_ARGUMENT, "allocator callbacks are required");
    if (!out_color || !out_depth) return fail(GSB_ERR_INVALID_ARGUMENT, "out_color / out_depth are required");
    cudaStream_t s = (cudaStream_t)stream;
    const FwdParams p = make_params(args);
    const GeomLayout GL = GeomLayout::make(p.P);
    const ImageLayout IL = ImageLayout::make(p.W, p.H);
    char* geom = (char*)geometry_alloc(geometry_user, GL.total);
    char* image = (char*)image_alloc(image_user, IL.total);
    if (!geom || !image) return fail(GSB_ERR_WORKSPACE, "geometry / image allocator returned NULL");
    if (int rc = forward_stage1(p, geom, GL, image, IL, radii, 0xffffffffu, s)) return rc;
    // the one synchronisation of the drop-in path (rasterizer_impl.cu:285 does a blocking cudaMemcpy)
    GeomHeader h;
    GSB_CUDA_CHECK(cudaMemcpyAsync(&h, geom + GL.header, sizeof(uint32_t) * 8, cudaMemcpyDeviceToHost, s));
    GSB_CUDA_CHECK(cudaStreamSynchronize(s));
    const long long R = h.num_rendered;
    if (R >= (1ll << 30)) return fail(GSB_ERR_OVERFLOW, "num_rendered %lld exceeds the 2^30 instance limit", R);
    const BinningLayout BL = BinningLayout::make(R);
    char* binning = (char*)binning_alloc(binning_user, BL.total);
    if (!binning) return fail(GSB_ERR_WORKSPACE, "binning allocator returned NULL");
    if (int rc = forward_stage2(p, geom, GL, binning, BL, image, IL, R, out_color, out_depth, nullptr, s)) return rc;
    return (int)R;
}

static int forward_ws_impl(const gsb_raster_args* args, void* geometry, size_t geometry_bytes, void* binning, size_t binning_bytes,
                           long long max_rendered, void* image, size_t image_bytes, float* out_color, float* out_depth,
                           float* out_depth_sil, int* radii, gsb_stream_t stream)
{
    if (int rc = validate(args, true)) return rc;
    if (!out_color || !out_depth) return fail(GSB_ERR_INVALID_ARGUMENT, "out_color / out_depth are required");
    if (max_rendered < 0 || max_rendered >= (1ll << 30)) return fail(GSB_ERR_INVALID_ARGUMENT, "max_rendered out of range");
    const FwdParams p = make_params(args);
    const GeomLayout GL = GeomLayout::make(p.P);
    const ImageLayout IL = ImageLayout::make(p.W, p.H);
    const BinningLayout BL = BinningLayout::make(max_rendered);
    if (!geometry || geometry_bytes < GL.total) return fail(GSB_ERR_WORKSPACE, "geometry workspace too small (%zu < %zu)", geometry_bytes, GL.total);
    if (!image || image_bytes < IL.total) return fail(GSB_ERR_WORKSPACE, "image workspace too small (%zu < %zu)", image_bytes, IL.total);
    if (!binning || binning_bytes < BL.total) return fail(GSB_ERR_WORKSPACE, "binning workspace too small (%zu < %zu)", binning_bytes, BL.total);
    cudaStream_t s = (cudaStream_t)stream;
    if (int rc = forward_stage1(p, (char*)geometry, GL, (char*)image, IL, radii, (uint32_t)max_rendered, s)) return rc;
    return forward_stage2(p, (char*)geometry, GL, (char*)binning, BL, (char*)image, IL, max_rendered, out_color, out_depth,
                          out_depth_sil, s);
}

int gsb_forward_ws(const gsb_raster_args* args, void* geometry, size_t geometry_bytes, void* binning, size_t binning_bytes,
                   long long max_rendered, void* image, size_t image_bytes, float* out_color, float* out_depth, int* radii,
                   gsb_stream_t stream)
{
    return forward_ws_impl(args, geometry, geometry_bytes, binning, binning_bytes, max_rendered, image, image_bytes, out_color,
                           out_depth, nullptr, radii, stream);
}

int gsb_forward_fused_ws(const gsb_raster_args* args, void* geometry, size_t geometry_bytes, void* binning, size_t binning_bytes,
                         long long max_rendered, void* image, size_t image_bytes, float* out_color, float* out_depth_sil,
                         float* out_median_depth, int* radii, gsb_stream_t stream)
{
    if (!out_depth_sil) return fail(GSB_ERR_INVALID_ARGUMENT, "out_depth_sil is required");
    return forward_ws_impl(args, geometry, geometry_bytes, binning, binning_bytes, max_rendered, image, image_bytes, out_color,
                           out_median_depth, out_depth_sil, radii, stream);
}

long long gsb_num_rendered(const void* geometry, gsb_stream_t stream)
{
    if (!geometry) return fail(GSB_ERR_INVALID_ARGUMENT, "geometry is NULL");
    GeomHeader h;
    cudaStream_t s = (cudaStream_t)stream;
    GSB_CUDA_CHECK(cudaMemcpyAsync(&h, geometry, sizeof(uint32_t) * 8, cudaMemcpyDeviceToHost, s));
    GSB_CUDA_CHECK(cudaStreamSynchronize(s));
    if (h.magic != GEOM_MAGIC) return fail(GSB_ERR_INVALID_ARGUMENT, "geometry blob has no valid header");
    if (h.overflow) return fail(GSB_ERR_OVERFLOW, "num_rendered %u exceeded the binning capacity %u", h.num_rendered, h.capacity);
    return (long long)h.num_rendered;
}

static int backward_impl(const gsb_raster_args* args, const int* radii, const void* geometry, const void* binning,
                         const void* image, const float* dL_dpix, const float* dL_ddepth_sil, const gsb_grad_outputs* grads,
                         float* dL_dzcolor, int z_attached, gsb_stream_t stream)
{
    if (int rc = validate(args, true, false)) return rc;  // opacities are not an input of the backward (rasterizer.h:55-83)
    if (!geometry || !binning || !image) return fail(GSB_ERR_INVALID_ARGUMENT, "forward state blobs are required");
    if (!dL_dpix || !grads) return fail(GSB_ERR_INVALID_ARGUMENT, "dL_dpix / grads are required");
    cudaStream_t s = (cudaStream_t)stream;
    const FwdParams p = make_params(args);
    const GeomLayout GL = GeomLayout::make(p.P);
    const ImageLayout IL = ImageLayout::make(p.W, p.H);
    char* geom = (char*)const_cast<void*>(geometry);  // the packed accumulators live in the blob
    if (int rc = launch_blend_backward(p, geom, GL, (const char*)binning, (const char*)image, IL, dL_dpix, dL_ddepth_sil, s))
        return rc;
    return launch_gauss_backward(p, geom, GL, radii, *grads, dL_dzcolor, z_attached, s);
}

int gsb_backward(const gsb_raster_args* args, long long R, const int* radii, const void* geometry, const void* binning,
                 const void* image, const float* dL_dpix, const gsb_grad_outputs* grads, gsb_stream_t stream)
{
    (void)R;  // tile ranges in the image blob already bound every list
    return backward_impl(args, radii, geometry, binning, image, dL_dpix, nullptr, grads, nullptr, 0, stream);
}

int gsb_backward_fused(const gsb_raster_args* args, const int* radii, const void* geometry, const void* binning,
                       const void* image, const float* dL_dcolor, const float* dL_ddepth_sil, const gsb_grad_outputs* grads,
                       float* dL_dzcolor, int z_attached, gsb_stream_t stream)
{
    if (!dL_ddepth_sil) return fail(GSB_ERR_INVALID_ARGUMENT, "dL_ddepth_sil is required");
    return backward_impl(args, radii, geometry, binning, image, dL_dcolor, dL_ddepth_sil, grads, dL_dzcolor, z_attached, stream);
}

int gsb_backward_fused_update(const gsb_raster_args* args, const int* radii, const void* geometry, const void* binning,
                              const void* image, const float* dL_dcolor, const float* dL_ddepth_sil, int z_attached,
                              const gsb_map_update* u, gsb_stream_t stream)
{
    if (int rc = validate(args, true, false)) return rc;
    if (!geometry || !binning || !image) return fail(GSB_ERR_INVALID_ARGUMENT, "forward state blobs are required");
    if (!dL_dcolor || !u) return fail(GSB_ERR_INVALID_ARGUMENT, "dL_dcolor / update are required");
    if (args->shs || args->cov3D_precomp || !args->colors_precomp)
        return fail(GSB_ERR_INVALID_ARGUMENT, "backward_fused_update: precomputed colours, scales and rotations only");
    const FwdParams p = make_params(args);
    if (p.band_y0 > 0 || p.band_y1 < p.tiles_y) return fail(GSB_ERR_INVALID_ARGUMENT, "backward_fused_update: whole image only");
    int ngrads = 0;
    for (int g = 0; g < 5; g++) {
        if (p.P > 0 && (!u->params[g] || !u->exp_avg[g] || !u->exp_avg_sq[g]))
            return fail(GSB_ERR_INVALID_ARGUMENT, "backward_fused_update: parameters and both Adam moments are required");
        ngrads += u->grads[g] ? 1 : 0;
    }
    if (ngrads != 0 && ngrads != 5) return fail(GSB_ERR_INVALID_ARGUMENT, "backward_fused_update: grads are all NULL or all set");
    if (p.P > 0 && args->colors_precomp != u->params[1])
        return fail(GSB_ERR_INVALID_ARGUMENT, "backward_fused_update: colors_precomp must be the rgb parameter group");
    if (!u->Tcw || u->step < 1 || (u->max_scalar > 0.f && !u->reg_terms))
        return fail(GSB_ERR_INVALID_ARGUMENT, "backward_fused_update: Tcw, step >= 1 and (with regularisers) reg_terms are required");
    cudaStream_t s = (cudaStream_t)stream;
    const GeomLayout GL = GeomLayout::make(p.P);
    const ImageLayout IL = ImageLayout::make(p.W, p.H);
    char* geom = (char*)const_cast<void*>(geometry);
    if (int rc = launch_blend_backward(p, geom, GL, (const char*)binning, (const char*)image, IL, dL_dcolor, dL_ddepth_sil, s)) return rc;
    return launch_map_update(p, geom, GL, radii, z_attached, *u, s);
}

int gsb_backward_fused_pose(const gsb_raster_args* args, const int* radii, const void* geometry, const void* binning,
                            const void* image, const float* dL_dcolor, const float* dL_ddepth_sil, int z_attached,
                            const float* means_world, float* dL_dTcw, gsb_stream_t stream)
{
    if (int rc = validate(args, true, false)) return rc;
    if (!geometry || !binning || !image) return fail(GSB_ERR_INVALID_ARGUMENT, "forward state blobs are required");
    if (!dL_dcolor || !dL_dTcw || (args->P > 0 && !means_world))
        return fail(GSB_ERR_INVALID_ARGUMENT, "backward_fused_pose: dL_dcolor, means_world and dL_dTcw are required");
    if (args->shs || args->cov3D_precomp)
        return fail(GSB_ERR_INVALID_ARGUMENT, "backward_fused_pose: precomputed colours, scales and rotations only");
    const FwdParams p = make_params(args);
    cudaStream_t s = (cudaStream_t)stream;
    const GeomLayout GL = GeomLayout::make(p.P);
    const ImageLayout IL = ImageLayout::make(p.W, p.H);
    char* geom = (char*)const_cast<void*>(geometry);
    if (int rc = launch_blend_backward(p, geom, GL, (const char*)binning, (const char*)image, IL, dL_dcolor, dL_ddepth_sil, s)) return rc;
    return launch_pose_gradient(p, geom, GL, radii, z_attached, means_world, dL_dTcw, s);
}

int gsb_visible_filter(const gsb_raster_args* args, int* radii, gsb_stream_t stream)
{
    if (int rc = validate(args, false)) return rc;
    if (!radii && args->P > 0) return fail(GSB_ERR_INVALID_ARGUMENT, "radii is required");
    return launch_visible_filter(make_params(args), radii, (cudaStream_t)stream);
}

int gsb_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix, uint8_t* present,
                     gsb_stream_t stream)
{
    (void)projmatrix;  // the reference's in_frustum only tests view-space z (auxiliary.h:139-163)
    if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) return fail(GSB_ERR_INVALID_ARGUMENT, "mark_visible: bad arguments");
    return launch_mark_visible(P, means3D, viewmatrix, present, (cudaStream_t)stream);
}

size_t gsb_knn_workspace_bytes(int P) { return knn_workspace_bytes(P); }
int gsb_knn_mean_dist2(int P, const float* points, float* mean_dist2, void* workspace, size_t workspace_bytes,
                       gsb_stream_t stream)
{
    if (P < 0 || (P > 0 && (!points || !mean_dist2))) return fail(GSB_ERR_INVALID_ARGUMENT, "knn: bad arguments");
    if (P == 0) return GSB_OK;
    if (!workspace || workspace_bytes < knn_workspace_bytes(P)) return fail(GSB_ERR_WORKSPACE, "knn workspace too small");
    return launch_knn(P, points, mean_dist2, (char*)workspace, (cudaStream_t)stream);
}

int gsb_prologue(int P, const float* Tcw, const float* means_world, const float* logit_opacities, const float* unnorm_quats,
                 const float* log_scales, float* means_cam, float* opacities, float* rotations, float* scales,
                 gsb_stream_t stream)
{
    if (P < 0) return fail(GSB_ERR_INVALID_ARGUMENT, "prologue: P < 0");
    if (P > 0 && ((means_cam && (!Tcw || !means_world)) || (opacities && !logit_opacities) || (rotations && !unnorm_quats) ||
                  (scales && !log_scales)))
        return fail(GSB_ERR_INVALID_ARGUMENT, "prologue: an output was requested without its input");
    return launch_prologue(P, Tcw, means_world, logit_opacities, unnorm_quats, log_scales, means_cam, opacities, rotations,
                           scales, (cudaStream_t)stream);
}

int gsb_prologue_backward(int P, const float* Tcw, const float* means_world, const float* logit_opacities,
                          const float* unnorm_quats, const float* log_scales, const float* dL_dmeans_cam,
                          const float* dL_dopacities, const float* dL_drotations, const float* dL_dscales,
                          float* dL_dmeans_world, float* dL_dlogit_opacities, float* dL_dunnorm_quats, float* dL_dlog_scales,
                          float* dL_dTcw, gsb_stream_t stream)
{
    if (P < 0) return fail(GSB_ERR_INVALID_ARGUMENT, "prologue_backward: P < 0");
    if (P > 0 && (((dL_dmeans_world || dL_dTcw) && (!dL_dmeans_cam || !Tcw || !means_world)) ||
                  (dL_dlogit_opacities && (!dL_dopacities || !logit_opacities)) ||
                  (dL_dunnorm_quats && (!dL_drotations || !unnorm_quats)) || (dL_dlog_scales && (!dL_dscales || !log_scales))))
        return fail(GSB_ERR_INVALID_ARGUMENT, "prologue_backward: an output was requested without its inputs");
    return launch_prologue_backward(P, Tcw, means_world, logit_opacities, unnorm_quats, log_scales, dL_dmeans_cam, dL_dopacities,
                                    dL_drotations, dL_dscales, dL_dmeans_world, dL_dlogit_opacities, dL_dunnorm_quats,
                                    dL_dlog_scales, dL_dTcw, (cudaStream_t)stream);
}

int gsb_pose_grad(int P, const float* means_world, const float* dL_dmeans_cam, float* dL_dTcw, gsb_stream_t stream)
{
    if (P < 0 || !dL_dTcw || (P > 0 && (!means_world || !dL_dmeans_cam))) return fail(GSB_ERR_INVALID_ARGUMENT, "pose_grad: bad arguments");
    return launch_prologue_backward(P, nullptr, means_world, nullptr, nullptr, nullptr, dL_dmeans_cam, nullptr, nullptr, nullptr,
                                    nullptr, nullptr, nullptr, nullptr, dL_dTcw, (cudaStream_t)stream);
}

int gsb_adam_step(long long n, float* param, const float* grad, float* exp_avg, float* exp_avg_sq, double lr, double beta1,
                  double beta2, double eps, long long step, gsb_stream_t stream)
{
    if (n < 0 || step < 1 || (n > 0 && (!param || !grad || !exp_avg || !exp_avg_sq)))
        return fail(GSB_ERR_INVALID_ARGUMENT, "adam_step: bad arguments");
    return launch_adam(n, param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, step, (cudaStream_t)stream);
}

int gsb_adam_step_groups(int ngroups, const long long* group_sizes_host, const float* lrs_host, float* param, const float* grad,
                         float* exp_avg, float* exp_avg_sq, double beta1, double beta2, double eps, long long step, gsb_stream_t stream)
{
    if (ngroups < 0 || ngroups > 8 || step < 1 || (ngroups > 0 && (!group_sizes_host || !lrs_host || !param || !grad || !exp_avg || !exp_avg_sq)))
        return fail(GSB_ERR_INVALID_ARGUMENT, "adam_step_groups: need 0 <= ngroups <= 8, step >= 1 and all arrays");
    for (int g = 0; g < ngroups; g++)
        if (group_sizes_host[g] < 0) return fail(GSB_ERR_INVALID_ARGUMENT, "adam_step_groups: negative group size");
    return launch_adam_groups(ngroups, group_sizes_host, lrs_host, param, grad, exp_avg, exp_avg_sq, beta1, beta2, eps, step,
                              (cudaStream_t)stream);
}

int gsb_scale_regulariser(int P, const float* log_scales, float max_scalar, float w_scalar, float w_long, float* dL_dlog_scales,
                          float* terms, gsb_stream_t stream)
{
    if (P < 0 || !terms || (P > 0 && !log_scales)) return fail(GSB_ERR_INVALID_ARGUMENT, "scale_regulariser: bad arguments");
    // terms[4..7] of the caller's 8-float block double as the reduction scratch
    return launch_scale_regulariser(P, log_scales, max_scalar, w_scalar, w_long, dL_dlog_scales, terms, terms + 4, (cudaStream_t)stream);
}

// ---- host-buffer convenience ----------------------------------------------------------------
namespace {
struct HostLayout {
    size_t means, colors, shs, opac, scales, rots, cov, bg, view, proj, campos, dpix;
    size_t out_color, out_depth, radii;
    size_t g_mean2D, g_conic, g_opac, g_color, g_mean3D, g_cov3D, g_sh, g_scale, g_rot;
    size_t geom, image, binning, total;
    GeomLayout GL;
    ImageLayout IL;
    BinningLayout BL;
    static HostLayout make(int P, int M, int W, int H, long long max_rendered)
    {
        HostLayout L;
        size_t off = 0;
        auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
        const size_t Pz = (size_t)(P > 0 ? P : 0), HW = (size_t)W * H, Mz = (size_t)(M > 0 ? M : 0);
        L.means = take(Pz * 12); L.colors = take(Pz * 12); L.shs = take(Pz * Mz * 12); L.opac = take(Pz * 4);
        L.scales = take(Pz * 12); L.rots = take(Pz * 16); L.cov = take(Pz * 24); L.bg = take(12); L.view = take(64);
        L.proj = take(64); L.campos = take(12); L.dpix = take(HW * 12);
        L.out_color = take(HW * 12); L.out_depth = take(HW * 4); L.radii = take(Pz * 4);
        L.g_mean2D = take(Pz * 12); L.g_conic = take(Pz * 16); L.g_opac = take(Pz * 4); L.g_color = take(Pz * 12);
        L.g_mean3D = take(Pz * 12); L.g_cov3D = take(Pz * 24); L.g_sh = take(Pz * Mz * 12); L.g_scale = take(Pz * 12);
        L.g_rot = take(Pz * 16);
        L.GL = GeomLayout::make(P); L.IL = ImageLayout::make(W, H); L.BL = BinningLayout::make(max_rendered);
        L.geom = take(L.GL.total); L.image = take(L.IL.total); L.binning = take(L.BL.total);
        L.total = off;
        return L;
    }
};
}  // namespace

size_t gsb_host_scratch_bytes(int P, int M, int width, int height, long long max_rendered)
{
    if (width <= 0 || height <= 0) return 0;
    return HostLayout::make(P, M, width, height, max_rendered).total;
}

// Shared body of the two host-buffer entry points.  Three streams per frame: ALL uploads go through one per-thread upload
// stream, ALL downloads through one per-thread download stream (a stream that alternates directions gets little of the
// link's duplex bandwidth: tools/pcie_pipe_probe.py measured 2.0 ms per 60 + 65 MB step that way against 1.36 ms with one
// stream per direction), the kernels run on the caller's `stream`, events chain the three.  Inside one frame the upload of
// dL/dpixel runs under the forward pass and the download of image / depth / radii under the backward pass; across frames in
// flight (gsb_forward_backward_host_async with K scratch sets and K streams) frame i's gradient download runs under
// frame i+1's kernels and frame i+2's upload.  `stream` completes only after the frame's last download.
static int host_frame(const gsb_raster_args* host_args, long long max_rendered, const float* dL_dpix_host, float* out_color_host,
                      float* out_depth_host, int* radii_host, const gsb_grad_outputs* host_grads, void* device_scratch,
                      size_t device_scratch_bytes, unsigned int* status_host, gsb_stream_t stream)
{
    if (int rc = validate_host(host_args)) return rc;
    if (max_rendered < 0 || max_rendered >= (1ll << 30)) return fail(GSB_ERR_INVALID_ARGUMENT, "max_rendered out of range");
    const gsb_raster_args& h = *host_args;
    const int P = h.P, M = h.shs ? h.M : 0, W = h.width, H = h.height;
    const HostLayout L = HostLayout::make(P, M, W, H, max_rendered);
    if (!device_scratch || device_scratch_bytes < L.total)
        return fail(GSB_ERR_WORKSPACE, "device scratch too small (%zu < %zu)", device_scratch_bytes, L.total);
    cudaStream_t s = (cudaStream_t)stream;
    struct HostStreams { cudaStream_t up = nullptr, dn = nullptr; cudaEvent_t ev_prev, ev_params, ev_dl, ev_fwd, ev_bwd, ev_dn; };
    static thread_local HostStreams streams[64];   // one set per device this thread has used (streams and events are per device)
    int device = 0;
    GSB_CUDA_CHECK(cudaGetDevice(&device));
    if (device < 0 || device >= 64) return fail(GSB_ERR_UNSUPPORTED, "device ordinal %d out of range", device);
    HostStreams& hs = streams[device];
    if (!hs.up) {
        GSB_CUDA_CHECK(cudaStreamCreateWithFlags(&hs.up, cudaStreamNonBlocking));
        GSB_CUDA_CHECK(cudaStreamCreateWithFlags(&hs.dn, cudaStreamNonBlocking));
        for (cudaEvent_t* e : {&hs.ev_prev, &hs.ev_params, &hs.ev_dl, &hs.ev_fwd, &hs.ev_bwd, &hs.ev_dn})
            GSB_CUDA_CHECK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    }
    char* d = (char*)device_scratch;
    const size_t Pz = (size_t)P, HW = (size_t)W * H;
    auto up = [&](size_t off, const void* src, size_t bytes) -> const float* {
        if (!src || !bytes) return nullptr;
        cudaMemcpyAsync(d + off, src, bytes, cudaMemcpyHostToDevice, hs.up);
        return reinterpret_cast<const float*>(d + off);
    };
    // an event is consumed by the cudaStreamWaitEvent issued right after its record (the wait snapshots the record), so the
    // per-thread events can be re-recorded by the next frame while this one is still in flight
    GSB_CUDA_CHECK(cudaEventRecord(hs.ev_prev, s));            // the scratch may still be in use by earlier work on `stream`
    GSB_CUDA_CHECK(cudaStreamWaitEvent(hs.up, hs.ev_prev, 0));
    gsb_raster_args a = h;
    a.means3D = up(L.means, h.means3D, Pz * 12);
    a.colors_precomp = up(L.colors, h.colors_precomp, Pz * 12);
    a.shs = up(L.shs, h.shs, Pz * M * 12);
    a.opacities = up(L.opac, h.opacities, Pz * 4);
    a.scales = up(L.scales, h.scales, Pz * 12);
    a.rotations = up(L.rots, h.rotations, Pz * 16);
    a.cov3D_precomp = up(L.cov, h.cov3D_precomp, Pz * 24);
    a.background = up(L.bg, h.background, 12);
    a.viewmatrix = up(L.view, h.viewmatrix, 64);
    a.projmatrix = up(L.proj, h.projmatrix, 64);
    a.cam_pos = up(L.campos, h.cam_pos, 12);
    GSB_CUDA_CHECK(cudaEventRecord(hs.ev_params, hs.up));
    const float* dpix = dL_dpix_host ? up(L.dpix, dL_dpix_host, HW * 12) : nullptr;
    GSB_CUDA_CHECK(cudaEventRecord(hs.ev_dl, hs.up));
    GSB_CUDA_CHECK(cudaGetLastError());
    float* oc = reinterpret_cast<float*>(d + L.out_color);
    float* od = reinterpret_cast<float*>(d + L.out_depth);
    int* rd = reinterpret_cast<int*>(d + L.radii);
    GSB_CUDA_CHECK(cudaStreamWaitEvent(s, hs.ev_params, 0));
    if (int rc = gsb_forward_ws(&a, d + L.geom, L.GL.total, d + L.binning, L.BL.total, max_rendered, d + L.image, L.IL.total, oc,
                                od, rd, stream))
        return rc;
    GSB_CUDA_CHECK(cudaEventRecord(hs.ev_fwd, s));
    GSB_CUDA_CHECK(cudaStreamWaitEvent(hs.dn, hs.ev_fwd, 0));
    if (out_color_host) cudaMemcpyAsync(out_color_host, oc, HW * 12, cudaMemcpyDeviceToHost, hs.dn);
    if (out_depth_host) cudaMemcpyAsync(out_depth_host, od, HW * 4, cudaMemcpyDeviceToHost, hs.dn);
    if (radii_host && P) cudaMemcpyAsync(radii_host, rd, Pz * 4, cudaMemcpyDeviceToHost, hs.dn);
    if (dpix && host_grads) {
        gsb_grad_outputs g;
        auto dev = [&](size_t off, const float* host) { return host ? reinterpret_cast<float*>(d + off) : nullptr; };
        g.dL_dmean2D = dev(L.g_mean2D, host_grads->dL_dmean2D); g.dL_dconic = dev(L.g_conic, host_grads->dL_dconic);
        g.dL_dopacity = dev(L.g_opac, host_grads->dL_dopacity); g.dL_dcolor = dev(L.g_color, host_grads->dL_dcolor);
        g.dL_dmean3D = dev(L.g_mean3D, host_grads->dL_dmean3D); g.dL_dcov3D = dev(L.g_cov3D, host_grads->dL_dcov3D);
        g.dL_dsh = M ? dev(L.g_sh, host_grads->dL_dsh) : nullptr; g.dL_dscale = a.scales ? dev(L.g_scale, host_grads->dL_dscale) : nullptr;
        g.dL_drot = a.rotations ? dev(L.g_rot, host_grads->dL_drot) : nullptr;
        GSB_CUDA_CHECK(cudaStreamWaitEvent(s, hs.ev_dl, 0));
        if (int rc = gsb_backward(&a, -1, rd, d + L.geom, d + L.binning, d + L.image, dpix, &g, stream)) return rc;
        GSB_CUDA_CHECK(cudaEventRecord(hs.ev_bwd, s));
        GSB_CUDA_CHECK(cudaStreamWaitEvent(hs.dn, hs.ev_bwd, 0));
        auto down = [&](float* host, const float* devp, size_t bytes) {
            if (host && devp && bytes) cudaMemcpyAsync(host, devp, bytes, cudaMemcpyDeviceToHost, hs.dn);
        };
        down(host_grads->dL_dmean2D, g.dL_dmean2D, Pz * 12); down(host_grads->dL_dconic, g.dL_dconic, Pz * 16);
        down(host_grads->dL_dopacity, g.dL_dopacity, Pz * 4); down(host_grads->dL_dcolor, g.dL_dcolor, Pz * 12);
        down(host_grads->dL_dmean3D, g.dL_dmean3D, Pz * 12); down(host_grads->dL_dcov3D, g.dL_dcov3D, Pz * 24);
        down(host_grads->dL_dsh, g.dL_dsh, Pz * M * 12); down(host_grads->dL_dscale, g.dL_dscale, Pz * 12);
        down(host_grads->dL_drot, g.dL_drot, Pz * 16);
    }
    if (status_host)   // GeomHeader words {num_rendered, num_rendered_clamped, overflow}
        cudaMemcpyAsync(status_host, d + L.geom + offsetof(GeomHeader, num_rendered), 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, hs.dn);
    GSB_CUDA_CHECK(cudaEventRecord(hs.ev_dn, hs.dn));
    GSB_CUDA_CHECK(cudaStreamWaitEvent(s, hs.ev_dn, 0));   // `stream` completes only after the frame's downloads
    GSB_CUDA_CHECK(cudaGetLastError());
    return GSB_OK;
}

long long gsb_forward_backward_host(const gsb_raster_args* host_args, long long max_rendered, const float* dL_dpix_host,
                                    float* out_color_host, float* out_depth_host, int* radii_host,
                                    const gsb_grad_outputs* host_grads, void* device_scratch, size_t device_scratch_bytes,
                                    gsb_stream_t stream)
{
    if (int rc = host_frame(host_args, max_rendered, dL_dpix_host, out_color_host, out_depth_host, radii_host, host_grads,
                            device_scratch, device_scratch_bytes, nullptr, stream))
        return rc;
    const HostLayout L = HostLayout::make(host_args->P, host_args->shs ? host_args->M : 0, host_args->width, host_args->height, max_rendered);
    return gsb_num_rendered((char*)device_scratch + L.geom, stream);
}

int gsb_forward_backward_host_async(const gsb_raster_args* host_args, long long max_rendered, const float* dL_dpix_host,
                                    float* out_color_host, float* out_depth_host, int* radii_host,
                                    const gsb_grad_outputs* host_grads, void* device_scratch, size_t device_scratch_bytes,
                                    unsigned int* status_host, gsb_stream_t stream)
{
    if (!status_host) return fail(GSB_ERR_INVALID_ARGUMENT, "status_host (3 x uint32 of pinned host memory) is required");
    return host_frame(host_args, max_rendered, dL_dpix_host, out_color_host, out_depth_host, radii_host, host_grads, device_scratch,
                      device_scratch_bytes, status_host, stream);
}

// ---- introspection ----------------------------------------------------------------------------
int gsb_debug_image_state(const void* image, int width, int height, float* final_T, uint32_t* n_contrib, uint32_t* ranges,
                          gsb_stream_t stream)
{
    if (!image || width <= 0 || height <= 0) return fail(GSB_ERR_INVALID_ARGUMENT, "debug_image_state: bad arguments");
    const ImageLayout IL = ImageLayout::make(width, height);
    cudaStream_t s = (cudaStream_t)stream;
    const char* im = (const char*)image;
    const size_t HW = (size_t)width * height, T = (size_t)IL.tiles_x * IL.tiles_y;
    if (final_T) GSB_CUDA_CHECK(cudaMemcpyAsync(final_T, im + IL.final_T, HW * 4, cudaMemcpyDeviceToDevice, s));
    if (n_contrib) GSB_CUDA_CHECK(cudaMemcpyAsync(n_contrib, im + IL.n_contrib, HW * 4, cudaMemcpyDeviceToDevice, s));
    if (ranges) GSB_CUDA_CHECK(cudaMemcpyAsync(ranges, im + IL.ranges, T * 8, cudaMemcpyDeviceToDevice, s));
    return GSB_OK;
}

int gsb_debug_blended_pairs(const void* geometry, const void* binning, const void* image, int width, int height,
                            unsigned long long* count, gsb_stream_t stream)
{
    if (!geometry || !binning || !image || !count || width <= 0 || height <= 0)
        return fail(GSB_ERR_INVALID_ARGUMENT, "debug_blended_pairs: bad arguments");
    const ImageLayout IL = ImageLayout::make(width, height);
    cudaStream_t s = (cudaStream_t)stream;
    GSB_CUDA_CHECK(cudaMemsetAsync(count, 0, sizeof(unsigned long long), s));
    const char* im = (const char*)image;
    count_blended_kernel<<<dim3(IL.tiles_x, IL.tiles_y), BLEND_THREADS, 0, s>>>(
        reinterpret_cast<const uint2*>(im + IL.ranges), reinterpret_cast<const uint32_t*>(im + IL.n_contrib),
        reinterpret_cast<const uint32_t*>(im + IL.hits_tail), (const char*)binning, reinterpret_cast<const GeomHeader*>(geometry),
        width, height, count);
    GSB_LAUNCH_CHECK();
    return GSB_OK;
}

int gsb_debug_binning_state(const void* geometry, const void* binning, long long R, uint32_t* point_list, gsb_stream_t stream)
{
    (void)geometry;
    if (!binning || R < 0) return fail(GSB_ERR_INVALID_ARGUMENT, "debug_binning_state: bad arguments");
    if (point_list && R)
        GSB_CUDA_CHECK(cudaMemcpyAsync(point_list, binning, (size_t)R * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return GSB_OK;
}

int gsb_debug_geometry_state(const void* geometry, int P, float* depths, float* means2D, float* conic_opacity,
                             uint32_t* tiles_touched, gsb_stream_t stream)
{
    if (!geometry || P < 0) return fail(GSB_ERR_INVALID_ARGUMENT, "debug_geometry_state: bad arguments");
    if (P == 0) return GSB_OK;
    const GeomLayout GL = GeomLayout::make(P);
    const char* g = (const char*)geometry;
    unpack_geometry_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        P, reinterpret_cast<const SplatRec*>(g + GL.rec), reinterpret_cast<const int*>(g + GL.radii),
        reinterpret_cast<const uint32_t*>(g + GL.tiles_touched), depths, means2D, conic_opacity, tiles_touched);
    GSB_LAUNCH_CHECK();
    return GSB_OK;
}

}  // extern "C"
