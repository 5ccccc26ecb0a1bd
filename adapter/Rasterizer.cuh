// adapter/Rasterizer.cuh -- drop-in replacement of GSORB-SLAM's include/Rasterizer.cuh.
//
// Keeps every name Render.cc / Gaussian.cc use (namespace ORB_SLAM2: GaussianRasterizationSettings,
// GaussianRasterizer::{forward, Visable, mark_visible}, rasterize_gaussians, filter_radii,
// RasterizeGaussiansCUDA, RasterizeGaussiansBackwardCUDA, RasterizeGaussiansfilterCUDA, markVisible,
// _RasterizeGaussians) with the same argument order, defaults, return values and exception types
// (reference: include/Rasterizer.cuh:28-125, 127-282, 284-380), but forwards to the C ABI of
// libgsb.so (include/gsb.h) instead of libCudaRasterizer.so.  libtorch tensors exist at this
// boundary only; nothing below it sees a torch type.
//
// Differences a maintainer should know (none changes a result):
//   * the seven scalar settings are passed as plain values, not as 0-dim CUDA tensors that are
//     read back with .item() (7 device->host syncs per call in the reference, Rasterizer.cuh:151-157);
//   * scratch blobs are NOT zero-filled (the kernels write everything they read);
//   * gradient tensors are torch::empty (fully written by gsb_backward) instead of 9 torch::zeros;
//   * the stream is at::cuda::getCurrentCUDAStream(), not the legacy default stream.
#pragma once
#include <torch/torch.h>

#include <functional>
#include <stdexcept>
#include <tuple>

#include "gsb.h"

namespace ORB_SLAM2 {

std::tuple<int, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
RasterizeGaussiansCUDA(const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& colors,
                       const torch::Tensor& opacity, const torch::Tensor& scales, const torch::Tensor& rotations,
                       const float scale_modifier, const torch::Tensor& cov3D_precomp, const torch::Tensor& viewmatrix,
                       const torch::Tensor& projmatrix, const float tan_fovx, const float tan_fovy, const int image_height,
                       const int image_width, const torch::Tensor& sh, const int degree, const torch::Tensor& campos,
                       const bool prefiltered, const int device_num);

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
RasterizeGaussiansBackwardCUDA(const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& radii,
                               const torch::Tensor& colors, const torch::Tensor& scales, const torch::Tensor& rotations,
                               const float scale_modifier, const torch::Tensor& cov3D_precomp, const torch::Tensor& viewmatrix,
                               const torch::Tensor& projmatrix, const float tan_fovx, const float tan_fovy,
                               const torch::Tensor& dL_dout_color, const torch::Tensor& sh, const int degree,
                               const torch::Tensor& campos, const torch::Tensor& geomBuffer, const int R,
                               const torch::Tensor& binningBuffer, const torch::Tensor& imageBuffer);

torch::Tensor markVisible(torch::Tensor& means3D, torch::Tensor& viewmatrix, torch::Tensor& projmatrix);

struct GaussianRasterizationSettings {
    int image_height;
    int image_width;
    float tanfovx;
    float tanfovy;
    torch::Tensor bg;
    float scale_modifier;
    torch::Tensor viewmatrix;
    torch::Tensor projmatrix;
    int sh_degree;
    torch::Tensor camera_center;
    bool prefiltered;
};

torch::Tensor filter_radii(torch::Tensor means3D, torch::Tensor scales, torch::Tensor rotations, int device_num,
                           GaussianRasterizationSettings raster_settings);

torch::Tensor RasterizeGaussiansfilterCUDA(const torch::Tensor& means3D, const torch::Tensor& scales,
                                           const torch::Tensor& rotations, const float scale_modifier,
                                           const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix,
                                           const float tan_fovx, const float tan_fovy, const int image_height,
                                           const int image_width, const bool prefiltered, int device_num);

torch::autograd::tensor_list rasterize_gaussians(torch::Tensor means3D, torch::Tensor means2D, torch::Tensor sh,
                                                 torch::Tensor colors_precomp, torch::Tensor opacities, torch::Tensor scales,
                                                 torch::Tensor rotations, torch::Tensor cov3Ds_precomp, int device_num,
                                                 GaussianRasterizationSettings raster_settings);

// Autograd node.  Unlike the reference (20 tensor arguments, 7 of them scalars wrapped in CUDA tensors)
// the settings travel as a struct; the differentiable inputs and the order of the returned gradients
// are unchanged (Rasterizer.cuh:259-266).
class _RasterizeGaussians : public torch::autograd::Function<_RasterizeGaussians> {
public:
    static torch::autograd::tensor_list forward(torch::autograd::AutogradContext* ctx, torch::Tensor means3D,
                                                torch::Tensor means2D, torch::Tensor sh, torch::Tensor colors_precomp,
                                                torch::Tensor opacities, torch::Tensor scales, torch::Tensor rotations,
                                                torch::Tensor cov3Ds_precomp, GaussianRasterizationSettings settings,
                                                int device_num);
    static torch::autograd::tensor_list backward(torch::autograd::AutogradContext* ctx,
                                                 torch::autograd::tensor_list grad_outputs);
};

class GaussianRasterizer : torch::nn::Module {
public:
    GaussianRasterizer() {}
    GaussianRasterizer(GaussianRasterizationSettings raster_settings) : raster_settings_(raster_settings) {}

    torch::Tensor mark_visible(torch::Tensor positions)
    {
        torch::NoGradGuard no_grad;
        return markVisible(positions, raster_settings_.viewmatrix, raster_settings_.projmatrix);
    }

    std::tuple<torch::Tensor, torch::Tensor, torch::Tensor> forward(torch::Tensor means3D, torch::Tensor means2D,
                                                                    torch::Tensor opacities, torch::Tensor shs = torch::Tensor(),
                                                                    torch::Tensor colors_precomp = torch::Tensor(),
                                                                    torch::Tensor scales = torch::Tensor(),
                                                                    torch::Tensor rotations = torch::Tensor(),
                                                                    torch::Tensor cov3D_precomp = torch::Tensor(),
                                                                    int device_num = 0)
    {
        if (shs.defined() == colors_precomp.defined())
            throw std::invalid_argument("Please provide exactly one of either SHs or precomputed colors!");
        if (((scales.defined() || rotations.defined()) && cov3D_precomp.defined()) ||
            (!scales.defined() && !rotations.defined() && !cov3D_precomp.defined()))
            throw std::invalid_argument("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!");
        // absent optional inputs travel as 0-element tensors, as in the reference (Rasterizer.cuh:320-334): autograd's
        // Function::apply needs every tensor argument to have a device
        auto empty = [&](torch::Tensor& t) { if (!t.defined()) t = torch::empty({0}, means3D.options().requires_grad(false)); };
        empty(shs); empty(colors_precomp); empty(scales); empty(rotations); empty(cov3D_precomp);
        auto result = rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp,
                                          device_num, raster_settings_);
        return {result[0], result[1], result[2]};
    }

    std::tuple<torch::Tensor> Visable(torch::Tensor means3D, torch::Tensor opacities, torch::Tensor scales = torch::Tensor(),
                                      torch::Tensor rotations = torch::Tensor(), int device_num = 0)
    {
        (void)opacities;
        return filter_radii(means3D, scales, rotations, device_num, raster_settings_);
    }

public:
    GaussianRasterizationSettings raster_settings_;
};

// ---- optional fast path (not in the reference): one five-channel pass instead of two ---------------
// Render::RenderForFrame / RenderStartTraking rasterize the same geometry twice per iteration: colours
// [r,g,b] and colours [z_cam, 1, 0] (src/Render.cc:445-448, :1068-1071).  rasterize_gaussians_fused
// returns {color [3,H,W], depth_sil [2,H,W], median_depth [1,H,W], radii [P]} from ONE pass
// (gsb_forward_fused_ws) with the gradients of both passes summed in ONE backward.  z_attached: the
// z_cam colour is a function of means3D (mapping mode) -> its gradient is added to d(means3D).z;
// false reproduces the detached colour of tracking mode (src/Render.cc:957).
torch::autograd::tensor_list rasterize_gaussians_fused(torch::Tensor means3D, torch::Tensor colors_precomp,
                                                       torch::Tensor opacities, torch::Tensor scales, torch::Tensor rotations,
                                                       GaussianRasterizationSettings raster_settings, bool z_attached);

class _RasterizeGaussiansFused : public torch::autograd::Function<_RasterizeGaussiansFused> {
public:
    static torch::autograd::tensor_list forward(torch::autograd::AutogradContext* ctx, torch::Tensor means3D,
                                                torch::Tensor colors_precomp, torch::Tensor opacities, torch::Tensor scales,
                                                torch::Tensor rotations, GaussianRasterizationSettings settings, bool z_attached);
    static torch::autograd::tensor_list backward(torch::autograd::AutogradContext* ctx,
                                                 torch::autograd::tensor_list grad_outputs);
};

// include/spatial.h: mean squared distance to the 3 nearest neighbours.
torch::Tensor distCUDA2(const torch::Tensor& points, torch::Device device);

}  // namespace ORB_SLAM2
