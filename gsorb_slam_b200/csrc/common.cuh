// common.cuh -- shared device/host helpers of libgsb (sm_100a only).
//
// Floating-point contract: every operation that feeds a DECISION (cull, ceil, trunc, == 0,
// alpha / transmittance thresholds) is written with explicit __fmul_rn / __fadd_rn / fmaf
// intrinsics, which nvcc never re-associates or contracts, in the exact fp32 sequence the
// reference sources compile to (forward.cu:74-152,155-256,261-401; auxiliary.h).  Radii,
// tile rectangles, sort keys, conics, n_contrib and the blended image are therefore
// bit-identical to the reference kernels -- see DESIGN.md "Arithmetic contract".  Gradient
// arithmetic uses ordinary expressions (tolerance 1e-3, tests/).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/gsb.h"

namespace gsb {

constexpr int TILE_X = 16;   // config.h:16 (tile-rect membership is part of the semantics)
constexpr int TILE_Y = 16;   // config.h:17
constexpr int NUM_SMS = 148; // B200
constexpr int TILE_CTR_STRIDE = 64;  // words between per-tile atomic counters (256 B)
constexpr int BLEND_THREADS = TILE_X * TILE_Y;   // one blend CTA per tile, one thread per pixel
constexpr int HIT_PIXELS = TILE_X * TILE_Y;      // hit words per window of a tile's list: one per pixel
constexpr int HIT_WINDOW = 32;                   // list entries covered by one hit word
constexpr int BUCKET_CAP = 8192;                 // records per tile that preprocess writes straight into the tile's bucket (ImageLayout::buckets)

// ---------------------------------------------------------------------------------------
// Packed per-Gaussian splat record: 48 bytes, 16-byte aligned, gathered by the blend
// kernels with 3 x 16-byte asynchronous copies.
//   a = { x_pix, y_pix, half2(extent_x, extent_y), power_threshold }   <- all the cull pass reads
//   b = { conic.x, conic.y, conic.z, opacity }
//   c = { r, g, b, depth }
// power_threshold: the EXACT boundary of the alpha >= 1/255 test in power space: the smallest float p with
// fmul_rn(opacity, expf(p)) >= 1/255 (found per Gaussian in preprocess.cu).  The forward kernel still applies the
// reference's alpha test (the threshold only saves it the exponential); the backward kernel decides by it alone.
// extent_x/y: conservative half-extents (pixels) of the alpha >= 1/255 footprint's bbox.
// ---------------------------------------------------------------------------------------
struct alignas(16) SplatRec {
    float4 a, b, c;
};
static_assert(sizeof(SplatRec) == 48, "record must be 48 bytes");

// Packed backward accumulators (one per Gaussian, 64 bytes = two aligned sectors, fp32 vector REDs): raw moments of u = G dL/dalpha
// and the colour sums of w = alpha T (blend_bwd.cu); gauss_bwd.cu turns them into dL/dmean2D, dL/dconic, dL/dopacity, dL/dcolor.
// One float4 per lane of a 4-lane group, so that every lane issues ONE 16-byte RED per visit (slots marked - stay zero):
//   a = { S u dx,    S u dx^2, S w d_b, - }
//   b = { S u dx dy, S w d_r,  -,       - }
//   c = { S u dy,    S u dy^2, S w d_z (fused 5-channel pass only), - }
//   d = { S u,       S w d_g,  -,       - }
constexpr int ACC_FLOATS = 16;
struct alignas(16) GradAcc {
    float4 a, b, c, d;
};
static_assert(sizeof(GradAcc) == ACC_FLOATS * 4, "accumulator record is 64 bytes");

// ---- opaque state blobs -------------------------------------------------------------------
constexpr uint32_t GEOM_MAGIC = 0x67736231u;  // "gsb1"

struct GeomHeader {            // first 256 bytes of the geometry blob (device memory)
    uint32_t magic;
    int32_t P;
    uint32_t num_rendered;     // total tile instances (written by the scan kernel)
    uint32_t num_rendered_clamped;  // min(num_rendered, capacity): what was actually binned
    uint32_t overflow;         // 1 if num_rendered > capacity
    uint32_t capacity;         // binning capacity in instances
    uint32_t sort_tile_counter[8];  // dynamic tile ids, one per radix pass
    uint32_t visible;          // number of Gaussians with radii > 0 (diagnostics)
    uint32_t layout_capacity;  // capacity the binning blob was laid out for (BinningLayout::make argument)
    uint32_t max_tile_len;     // longest tile list (selects the tile-sort size classes that have work)
    uint32_t pad[64 - 17];
};
// The header is zeroed by the forward before the first kernel; the scan kernel fills it.
static_assert(sizeof(GeomHeader) == 256, "header is 256 bytes");

inline __host__ __device__ size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct GeomLayout {            // offsets inside the geometry blob
    size_t header, rec, radii, tiles_touched, clamped, acc, total;
    int num_blocks;            // preprocess blocks of 256 Gaussians
    __host__ __device__ static GeomLayout make(int P)
    {
        GeomLayout L;
        size_t off = 0;
        const size_t Pz = (size_t)(P > 0 ? P : 0);
        L.num_blocks = (int)((Pz + 255) / 256);
        auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
        L.header = take(sizeof(GeomHeader));
        L.rec = take(Pz * sizeof(SplatRec));
        L.radii = take(Pz * 4);
        L.tiles_touched = take(Pz * 4);
        L.clamped = take(Pz);
        L.acc = take(Pz * sizeof(GradAcc));
        L.total = off;
        return L;
    }
};

struct ImageLayout {
    size_t final_T, n_contrib, ranges, tile_max_contrib, tile_count, tile_cursor, hits_tail, buckets, total;
    int tiles_x, tiles_y;
    __host__ __device__ static ImageLayout make(int W, int H)
    {
        ImageLayout L;
        size_t off = 0;
        auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
        L.tiles_x = (W + TILE_X - 1) / TILE_X;
        L.tiles_y = (H + TILE_Y - 1) / TILE_Y;
        const size_t HW = (size_t)W * H, T = (size_t)L.tiles_x * L.tiles_y;
        L.final_T = take(HW * 4);
        L.n_contrib = take(HW * 4);
        L.ranges = take(T * 8);
        L.tile_max_contrib = take(T * 8);   // highest n_contrib of each tile (both words)
        // one counter per TILE_CTR_STRIDE words: adjacent tiles land in different L2 slices, so the
        // ~2 M atomics of a frame are not funnelled through the few slices a dense array maps to
        // (word 0 of a slot: the tile's instance count; the counting atomic of preprocess returns the record's slot in the tile's bucket)
        L.tile_count = take((T + 1) * 4 * TILE_CTR_STRIDE);   // + one slot: completion counter of the preprocess CTAs
        L.tile_cursor = take(T * 4 * TILE_CTR_STRIDE);   // next free slot of each tile's segment (only the fallback `duplicate` pass uses it)
        L.hits_tail = take(T * HIT_PIXELS * 4);          // hit words of each tile's last, partial window (blend kernels)
        // Per-tile buckets of BUCKET_CAP unsorted (depth bits << 32 | id) records, written by preprocess at the slot its counting atomic
        // returned; the tile sort reads a tile's records from here whenever the frame's longest list fits (GeomHeader::max_tile_len <=
        // BUCKET_CAP: no `duplicate` pass at all).  Never cleared: only the first `count` records of a bucket are read.  64 KB per tile.
        L.buckets = take(T * (size_t)BUCKET_CAP * 8);
        L.total = off;
        return L;
    }
};

constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;  // 4096 keys per CTA
constexpr int SORT_RADIX_BITS = 8;
constexpr int SORT_RADIX = 1 << SORT_RADIX_BITS;
constexpr int SORT_MAX_PASSES = 8;

struct BinningLayout {
    // point_list sits at offset 0 whatever the capacity (gsb_backward relies on it): the
    // depth-sorted Gaussian ids, tile after tile.  `pairs` holds the unsorted
    // (depth bits << 32 | id) records bucketed by tile.  `hits` holds, per full 32-entry window
    // of a tile's list and per PIXEL of the tile (thread order of the blend kernels: warp = 8x4
    // region, lane = row-major pixel of the region), the bit mask of the entries that were blended
    // into that pixel (written by the forward blend; the backward blend replays exactly those
    // pairs).  Window w of a tile whose list starts at s lives at row (s >> 5) + w (256 words),
    // which never collides with the next tile's windows; each tile's last, partial window is
    // kept in the image blob (ImageLayout::hits_tail).  32 B per instance.
    size_t point_list, hits, pairs, total;
    long long capacity;
    __host__ __device__ static BinningLayout make(long long cap)
    {
        BinningLayout L;
        size_t off = 0;
        auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
        if (cap < 1) cap = 1;
        L.capacity = cap;
        L.point_list = take((size_t)cap * 4 + 64);
        L.hits = take(((size_t)cap / HIT_WINDOW + 1) * HIT_PIXELS * 4);
        L.pairs = take((size_t)cap * 8);
        L.total = off;
        return L;
    }
};

// ---------------------------------------------------------------------------------------
// Exact-order fp32 helpers (see file header).
// ---------------------------------------------------------------------------------------
// nvcc's contraction of a0*b0 + a1*b1 + a2*b2: middle product plain, first fused, third fused.
__device__ __forceinline__ float nv3(float a0, float b0, float a1, float b1, float a2, float b2)
{
    return fmaf(a2, b2, fmaf(a0, b0, __fmul_rn(a1, b1)));
}
// auxiliary.h:58-77 transformPoint4x3 / 4x4, one output row.
__device__ __forceinline__ float xform_row(const float* __restrict__ m, int r, float x, float y, float z)
{
    return __fadd_rn(nv3(m[r], x, m[4 + r], y, m[8 + r], z), m[12 + r]);
}
// auxiliary.h:41-44 ndc2Pix: evaluated in double, rounded once.
__device__ __forceinline__ float ndc2pix(float v, int S)
{
    return (float)(fma((double)v + 1.0, (double)S, -1.0) * 0.5);  // DADD, DFMA, DMUL, F2F as the reference
}
// auxiliary.h:46-56 getRect (16x16 tiles: "/16" is an exact multiply by 0.0625).
__device__ __forceinline__ void get_rect(float px, float py, int max_radius, int gx, int gy,
                                         uint32_t& minx, uint32_t& miny, uint32_t& maxx, uint32_t& maxy)
{
    const float r = (float)max_radius;
    const int lx = (int)(__fmul_rn(__fsub_rn(px, r), 0.0625f));
    const int ly = (int)(__fmul_rn(__fsub_rn(py, r), 0.0625f));
    const int hx = (int)(__fmul_rn(__fsub_rn(__fadd_rn(__fadd_rn(px, r), 16.0f), 1.0f), 0.0625f));
    const int hy = (int)(__fmul_rn(__fsub_rn(__fadd_rn(__fadd_rn(py, r), 16.0f), 1.0f), 0.0625f));
    minx = min((uint32_t)gx, (uint32_t)max(0, lx));
    miny = min((uint32_t)gy, (uint32_t)max(0, ly));
    maxx = min((uint32_t)gx, (uint32_t)max(0, hx));
    maxy = min((uint32_t)gy, (uint32_t)max(0, hy));
}

// forward.cu:118-152 computeCov3D, operation order as compiled for the reference.
__device__ __forceinline__ void compute_cov3d(float s0, float s1, float s2, float mod,
                                              float r, float x, float y, float z, float* cov3D)
{
    const float sx = __fmul_rn(mod, s0), sy = __fmul_rn(mod, s1), sz = __fmul_rn(mod, s2);
    const float xz = __fmul_rn(x, z), rx = __fmul_rn(r, x), rz = __fmul_rn(r, z);
    const float yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
    const float A = fmaf(r, y, xz);
    const float B = fmaf(-r, y, xz);
    const float C = fmaf(y, z, -rx);
    const float D = fmaf(y, z, rx);
    const float E = fmaf(x, y, -rz);
    const float F = fmaf(x, y, rz);
    const float G = fmaf(x, x, yy);
    const float H = __fadd_rn(yy, zz);
    const float I = fmaf(x, x, zz);
    const float R00 = __fsub_rn(1.f, __fadd_rn(H, H)), R01 = __fadd_rn(E, E), R02 = __fadd_rn(A, A);
    const float R10 = __fadd_rn(F, F), R11 = __fsub_rn(1.f, __fadd_rn(I, I)), R12 = __fadd_rn(C, C);
    const float R20 = __fadd_rn(B, B), R21 = __fadd_rn(D, D), R22 = __fsub_rn(1.f, __fadd_rn(G, G));
    const float M00 = __fmul_rn(sx, R00), M01 = __fmul_rn(sy, R01), M02 = __fmul_rn(sz, R02);
    const float M10 = __fmul_rn(sx, R10), M11 = __fmul_rn(sy, R11), M12 = __fmul_rn(sz, R12);
    const float M20 = __fmul_rn(sx, R20), M21 = __fmul_rn(sy, R21), M22 = __fmul_rn(sz, R22);
    cov3D[0] = nv3(M00, M00, M01, M01, M02, M02);
    cov3D[1] = nv3(M00, M10, M01, M11, M02, M12);
    cov3D[2] = nv3(M00, M20, M01, M21, M02, M22);
    cov3D[3] = nv3(M10, M10, M11, M11, M12, M12);
    cov3D[4] = nv3(M10, M20, M11, M21, M12, M22);
    cov3D[5] = nv3(M20, M20, M21, M21, M22, M22);
}

struct Cov2DInter {
    float tx, ty, tz, txtz, tytz;
    float T00, T01, T02, T10, T11, T12;
};

// forward.cu:74-113 computeCov2D (returns cov.x, cov.y, cov.z including the +0.3 low-pass).
__device__ __forceinline__ void compute_cov2d(float x, float y, float z, float focal_x, float focal_y,
                                              float tan_fovx, float tan_fovy, const float* c,
                                              const float* __restrict__ v, float* cov, Cov2DInter* inter)
{
    float tx = xform_row(v, 0, x, y, z);
    float ty = xform_row(v, 1, x, y, z);
    const float tz = xform_row(v, 2, x, y, z);
    const float limx = __fmul_rn(tan_fovx, 1.3f), limy = __fmul_rn(tan_fovy, 1.3f);
    const float txtz = __fdiv_rn(tx, tz), tytz = __fdiv_rn(ty, tz);
    const float cx = fminf(fmaxf(txtz, -limx), limx);
    const float cy = fminf(fmaxf(tytz, -limy), limy);
    tx = __fmul_rn(cx, tz);
    ty = __fmul_rn(cy, tz);
    const float tz2 = __fmul_rn(tz, tz);
    const float J00 = __fdiv_rn(focal_x, tz);
    const float J02 = __fdiv_rn(__fmul_rn(-tx, focal_x), tz2);
    const float J11 = __fdiv_rn(focal_y, tz);
    const float J12 = __fdiv_rn(__fmul_rn(-ty, focal_y), tz2);
    const float T00 = fmaf(v[2], J02, __fmul_rn(v[0], J00));
    const float T01 = fmaf(v[6], J02, __fmul_rn(v[4], J00));
    const float T02 = fmaf(v[10], J02, __fmul_rn(v[8], J00));
    const float T10 = fmaf(v[2], J12, __fmul_rn(v[1], J11));
    const float T11 = fmaf(v[6], J12, __fmul_rn(v[5], J11));
    const float T12 = fmaf(v[10], J12, __fmul_rn(v[9], J11));
    // A = transpose(T) * transpose(Vrk)
    const float A00 = nv3(T00, c[0], T01, c[1], T02, c[2]);
    const float A10 = nv3(T00, c[1], T01, c[3], T02, c[4]);
    const float A20 = nv3(T00, c[2], T01, c[4], T02, c[5]);
    const float A01 = nv3(T10, c[0], T11, c[1], T12, c[2]);
    const float A11 = nv3(T10, c[1], T11, c[3], T12, c[4]);
    const float A21 = nv3(T10, c[2], T11, c[4], T12, c[5]);
    // cov = A * T
    const float c00 = nv3(A00, T00, A10, T01, A20, T02);
    const float c01 = nv3(A01, T00, A11, T01, A21, T02);
    const float c11 = nv3(A01, T10, A11, T11, A21, T12);
    cov[0] = __fadd_rn(c00, 0.3f);
    cov[1] = c01;
    cov[2] = __fadd_rn(c11, 0.3f);
    if (inter) {
        inter->tx = tx; inter->ty = ty; inter->tz = tz; inter->txtz = txtz; inter->tytz = tytz;
        inter->T00 = T00; inter->T01 = T01; inter->T02 = T02;
        inter->T10 = T10; inter->T11 = T11; inter->T12 = T12;
    }
}

// The (pixel, splat) falloff exponent, forward.cu:346 as compiled for the reference:
//   fma(fma(dx, cx*dx, (cz*dy)*dy), -0.5, -((cy*dx)*dy))
__device__ __forceinline__ float splat_power(float dx, float dy, float conx, float cony, float conz)
{
    const float q = fmaf(dx, __fmul_rn(conx, dx), __fmul_rn(__fmul_rn(conz, dy), dy));
    return fmaf(q, -0.5f, -__fmul_rn(__fmul_rn(cony, dx), dy));
}

// ---- small utilities ------------------------------------------------------------------
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

// SH constants (auxiliary.h:20-38)
__device__ const float SH_C0 = 0.28209479177387814f;
__device__ const float SH_C1 = 0.4886025119029199f;
__device__ const float SH_C2[] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                  -1.0925484305920792f, 0.5462742152960396f};
__device__ const float SH_C3[] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                  0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                  -0.5900435899266435f};

// ---- host-side error plumbing (api.cu) ---------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

// ---- per-stage device timing (gsb_profile_begin / gsb_profile_end) -------------------------
enum Stage {
    ST_MEMSET = 0, ST_PREPROCESS, ST_SCAN, ST_DUPLICATE, ST_SORT_HIST, ST_SORT_PASS, ST_TILE_SORT, ST_BLEND_FWD,
    ST_BLEND_BWD, ST_GAUSS_BWD, ST_OTHER, ST_COUNT
};
// Scoped: when profiling is enabled on this thread, records a CUDA event on `s` before and
// after the enclosed launches (events on the launching stream, so they bracket exactly them).
struct StageTimer {
    int stage;
    cudaStream_t s;
    void* ev;
    StageTimer(int stage, cudaStream_t s);
    ~StageTimer();
};

#define GSB_CUDA_CHECK(expr)                                                              \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            gsb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return GSB_ERR_CUDA;                                                          \
        }                                                                                 \
    } while (0)

// cudaFuncSetAttribute once per kernel and device (not on every launch); a benign race at worst sets it twice
#define GSB_SET_ATTR_ONCE(kernel, attr, value)                                            \
    do {                                                                                  \
        static bool _done[64] = {};                                                       \
        int _d = 0;                                                                       \
        cudaGetDevice(&_d);                                                               \
        if (_d >= 0 && _d < 64 && !_done[_d]) {                                           \
            cudaFuncSetAttribute(kernel, attr, value);                                    \
            _done[_d] = true;                                                             \
        }                                                                                 \
    } while (0)

#define GSB_LAUNCH_CHECK()                                                                \
    do {                                                                                  \
        gsb::count_launch();                                                              \
        cudaError_t _e = cudaGetLastError();                                              \
        if (_e != cudaSuccess) {                                                          \
            gsb::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return GSB_ERR_CUDA;                                                          \
        }                                                                                 \
    } while (0)

// ---- launch entry points implemented in the other translation units ----------------------
struct FwdParams {
    int P, D, M, W, H, tiles_x, tiles_y;
    int band_y0, band_y1;   // tile rows [band_y0, band_y1) are binned and blended (tile-row shard); 0, tiles_y = all
    const float* background;
    const float* means3D;
    const float* shs;
    const float* colors_precomp;
    const float* opacities;
    const float* scales;
    float scale_modifier;
    const float* rotations;
    const float* cov3D_precomp;
    const float* viewmatrix;
    const float* projmatrix;
    const float* cam_pos;
    float tan_fovx, tan_fovy, focal_x, focal_y;
};

int launch_preprocess(const FwdParams& p, char* geom, const GeomLayout& GL, char* image, const ImageLayout& IL,
                      int* radii_out, uint32_t capacity, cudaStream_t s);
int launch_tile_scan(char* geom, const GeomLayout& GL, char* image, const ImageLayout& IL, uint32_t capacity, int P,
                     cudaStream_t s);
int launch_visible_filter(const FwdParams& p, int* radii, cudaStream_t s);
int launch_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present, cudaStream_t s);
int launch_sort_pairs(GeomHeader* hdr, uint64_t* const kbuf[2], uint32_t* const vbuf[2], int start, int passes,
                      uint32_t* hist, uint32_t* lookback, int sort_tiles, cudaStream_t s);
// grid_instances: host-side upper bound on the instance count used to size the sort grid
// (the exact count lives in the header on the device).
int launch_binning(const FwdParams& p, char* geom, const GeomLayout& GL, char* binning, const BinningLayout& BL,
                   char* image, const ImageLayout& IL, long long grid_instances, cudaStream_t s);
int launch_blend_forward(const FwdParams& p, char* geom, const GeomLayout& GL, char* binning, const BinningLayout& BL,
                         char* image, const ImageLayout& IL, float* out_color, float* out_depth, float* out_depth_sil,
                         cudaStream_t s);
int launch_blend_backward(const FwdParams& p, char* geom, const GeomLayout& GL, const char* binning,
                          const char* image, const ImageLayout& IL, const float* dL_dpix, const float* dL_ddepth_sil,
                          cudaStream_t s);
int launch_blend_backward_twophase(const FwdParams& p, char* geom, const GeomLayout& GL, const char* binning,
                                   const char* image, const ImageLayout& IL, const float* dL_dpix, const float* dL_ddepth_sil,
                                   cudaStream_t s);   // tuning builds only
int launch_gauss_backward(const FwdParams& p, const char* geom, const GeomLayout& GL, const int* radii,
                          const gsb_grad_outputs& g, float* dL_dzcolor, int z_attached, cudaStream_t s);
int launch_prologue(int P, const float* Tcw, const float* means_world, const float* logit, const float* quats,
                    const float* log_scales, float* means_cam, float* opac, float* rot, float* scales, cudaStream_t s);
int launch_prologue_backward(int P, const float* Tcw, const float* means_world, const float* logit, const float* quats,
                             const float* log_scales, const float* g_means, const float* g_opac, const float* g_rot,
                             const float* g_scales, float* d_means, float* d_logit, float* d_quats, float* d_log_scales,
                             float* dTcw, cudaStream_t s);
int launch_adam(long long n, float* param, const float* grad, float* m, float* v, double lr, double beta1, double beta2,
                double eps, long long step, cudaStream_t s);
int launch_adam_groups(int ngroups, const long long* sizes, const float* lrs, float* param, const float* grad, float* m, float* v,
                       double beta1, double beta2, double eps, long long step, cudaStream_t s);
int launch_scale_regulariser(int P, const float* log_scales, float max_scalar, float w_scalar, float w_long, float* d_log_scales,
                             float* terms, float* acc, cudaStream_t s);
int launch_scale_regulariser_sum(int P, const float* log_scales, float max_scalar, float* acc, cudaStream_t s);
int launch_map_update(const FwdParams& p, const char* geom, const GeomLayout& GL, const int* radii, int z_attached,
                      const gsb_map_update& u, cudaStream_t s);
int launch_pose_gradient(const FwdParams& p, const char* geom, const GeomLayout& GL, const int* radii, int z_attached,
                         const float* means_world, float* dTcw, cudaStream_t s);
size_t knn_workspace_bytes(int P);
int launch_knn(int P, const float* points, float* mean_dist2, char* ws, cudaStream_t s);

}  // namespace gsb
