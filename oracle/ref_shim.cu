// oracle/ref_shim.cu -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A thin extern "C" shim (written for this repo) around the UNMODIFIED reference
// CUDA rasterizer.  The reference translation units are compiled in place from
// /root/reference by oracle/Makefile (forward.cu, backward.cu, rasterizer_impl.cu,
// simple_knn.cu); nothing from them is copied here.  The shim exists only so that
// tests/, bench.py's reference arm and tests/golden/make_golden.py can drive the
// reference kernels with raw device pointers through ctypes, the same way
// src/Rasterizer.cu:136-297 drives them through libtorch.
//
// Interfaces wrapped (reference file:line):
//   CudaRasterizer::Rasterizer::forward        rasterizer.h:31-53   (impl.cu:199-345)
//   CudaRasterizer::Rasterizer::backward       rasterizer.h:55-83   (impl.cu:405-498)
//   CudaRasterizer::Rasterizer::visible_filter rasterizer.h:85-100  (impl.cu:348-401)
//   CudaRasterizer::Rasterizer::markVisible    rasterizer.h:24-29   (impl.cu:142-154)
//   SimpleKNN::knn                             simple_knn.h:15-19   (simple_knn.cu:185-221)
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include <functional>
#include "rasterizer.h"
#include "simple_knn.h"

extern "C" {

// Caller-owned scratch: like resizeFunctional (src/Rasterizer.cu:127-134) the
// callback must hand back >= bytes of ZEROED device memory.
typedef char* (*ref_alloc_fn)(void* user, size_t bytes);

int ref_forward(ref_alloc_fn geom_alloc, void* geom_user,
                ref_alloc_fn bin_alloc, void* bin_user,
                ref_alloc_fn img_alloc, void* img_user,
                int P, int D, int M, const float* background, int width, int height,
                const float* means3D, const float* shs, const float* colors_precomp,
                const float* opacities, const float* scales, float scale_modifier,
                const float* rotations, const float* cov3D_precomp,
                const float* viewmatrix, const float* projmatrix, const float* cam_pos,
                float tan_fovx, float tan_fovy, int prefiltered,
                float* out_color, float* out_depth, int* radii)
{
    std::function<char*(size_t)> g = [=](size_t n) { return geom_alloc(geom_user, n); };
    std::function<char*(size_t)> b = [=](size_t n) { return bin_alloc(bin_user, n); };
    std::function<char*(size_t)> i = [=](size_t n) { return img_alloc(img_user, n); };
    return CudaRasterizer::Rasterizer::forward(g, b, i, P, D, M, background, width, height,
        means3D, shs, colors_precomp, opacities, scales, scale_modifier, rotations,
        cov3D_precomp, viewmatrix, projmatrix, cam_pos, tan_fovx, tan_fovy,
        prefiltered != 0, out_color, out_depth, radii);
}

void ref_backward(int P, int D, int M, int R, const float* background, int width, int height,
                  const float* means3D, const float* shs, const float* colors_precomp,
                  const float* scales, float scale_modifier, const float* rotations,
                  const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                  const float* campos, float tan_fovx, float tan_fovy, const int* radii,
                  char* geom_buffer, char* binning_buffer, char* image_buffer,
                  const float* dL_dpix, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                  float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
                  float* dL_dscale, float* dL_drot)
{
    CudaRasterizer::Rasterizer::backward(P, D, M, R, background, width, height, means3D, shs,
        colors_precomp, scales, scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix,
        campos, tan_fovx, tan_fovy, radii, geom_buffer, binning_buffer, image_buffer, dL_dpix,
        dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale,
        dL_drot);
}

void ref_visible_filter(ref_alloc_fn geom_alloc, void* geom_user,
                        ref_alloc_fn bin_alloc, void* bin_user,
                        ref_alloc_fn img_alloc, void* img_user,
                        int P, int M, int width, int height, const float* means3D,
                        const float* scales, float scale_modifier, const float* rotations,
                        const float* viewmatrix, const float* projmatrix,
                        float tan_fovx, float tan_fovy, int prefiltered, int* radii)
{
    std::function<char*(size_t)> g = [=](size_t n) { return geom_alloc(geom_user, n); };
    std::function<char*(size_t)> b = [=](size_t n) { return bin_alloc(bin_user, n); };
    std::function<char*(size_t)> i = [=](size_t n) { return img_alloc(img_user, n); };
    CudaRasterizer::Rasterizer::visible_filter(g, b, i, P, M, width, height, means3D, scales,
        scale_modifier, rotations, viewmatrix, projmatrix, tan_fovx, tan_fovy,
        prefiltered != 0, radii);
}

void ref_mark_visible(int P, float* means3D, float* viewmatrix, float* projmatrix,
                      unsigned char* present)
{
    CudaRasterizer::Rasterizer::markVisible(P, means3D, viewmatrix, projmatrix, (bool*)present);
}

void ref_knn(int P, float* points, float* mean_dists)
{
    SimpleKNN::knn(P, (float3*)points, mean_dists);
}

// Byte offsets of the fields inside the reference's opaque state blobs, so the
// parity harness can slice final_T / n_contrib / ranges / point_list out of them
// (layout: impl.cu:156-195, 128-byte aligned bump allocation rasterizer_impl.h:22-27).
static size_t bump(size_t& off, size_t bytes) {
    size_t a = (off + 127) & ~size_t(127);
    off = a + bytes;
    return a;
}
void ref_image_state_offsets(size_t n_pixels, size_t* accum_alpha, size_t* n_contrib, size_t* ranges)
{
    size_t off = 0;
    *accum_alpha = bump(off, n_pixels * 4);
    *n_contrib = bump(off, n_pixels * 4);
    *ranges = bump(off, n_pixels * 8);
}
void ref_binning_state_offsets(size_t R, size_t* point_list, size_t* point_list_unsorted,
                               size_t* keys, size_t* keys_unsorted)
{
    size_t off = 0;
    *point_list = bump(off, R * 4);
    *point_list_unsorted = bump(off, R * 4);
    *keys = bump(off, R * 8);
    *keys_unsorted = bump(off, R * 8);
}
void ref_geometry_state_offsets(size_t P, size_t* depths, size_t* clamped, size_t* internal_radii,
                                size_t* means2D, size_t* cov3D, size_t* conic_opacity, size_t* rgb,
                                size_t* tiles_touched)
{
    size_t off = 0;
    *depths = bump(off, P * 4);
    *clamped = bump(off, P * 3);
    *internal_radii = bump(off, P * 4);
    *means2D = bump(off, P * 8);
    *cov3D = bump(off, P * 24);
    *conic_opacity = bump(off, P * 16);
    *rgb = bump(off, P * 12);
    *tiles_touched = bump(off, P * 4);
}

int ref_sync(void) { return (int)cudaDeviceSynchronize(); }

}  // extern "C"
