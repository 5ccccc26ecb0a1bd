// adapter/Rasterizer.cc -- replacement of GSORB-SLAM's src/Rasterizer.cu + src/spatial.cu over libgsb.so.
// Plain C++ (no device code): marshals libtorch tensors into the POD structs of include/gsb.h.
// Reference behaviour followed: src/Rasterizer.cu:8-73 (rasterize_gaussians), :75-122 (filter_radii),
// :136-217 (RasterizeGaussiansCUDA), :220-297 (RasterizeGaussiansBackwardCUDA), :299-318 (markVisible),
// :322-383 (RasterizeGaussiansfilterCUDA); src/spatial.cu:15-27 (distCUDA2).
#include "Rasterizer.cuh"

#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

namespace ORB_SLAM2 {
namespace {

const float* fptr(const torch::Tensor& t)
{
    // an undefined / empty optional input is the NULL pointer of the C ABI (Rasterizer.cuh:320-334)
    return (t.defined() && t.numel() > 0) ? t.data_ptr<float>() : nullptr;
}
torch::Tensor f32c(const torch::Tensor& t)
{
    return (t.defined() && t.numel() > 0) ? t.to(torch::kFloat32).contiguous() : t;
}
void check(long long rc)
{
    if (rc >= 0) return;
    const std::string msg = gsb_last_error();
    if (rc == GSB_ERR_INVALID_ARGUMENT) throw std::invalid_argument(msg);
    throw std::runtime_error("libgsb: " + msg);
}
void* stream() { return (void*)c10::cuda::getCurrentCUDAStream().stream(); }

struct Marshalled {
    torch::Tensor bg, means3D, colors, opacity, scales, rotations, cov3D, view, proj, sh, campos;
    gsb_raster_args a;
};
Marshalled marshal(const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& colors,
                   const torch::Tensor& opacity, const torch::Tensor& scales, const torch::Tensor& rotations, float scale_modifier,
                   const torch::Tensor& cov3D_precomp, const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix,
                   float tan_fovx, float tan_fovy, int image_height, int image_width, const torch::Tensor& sh, int degree,
                   const torch::Tensor& campos, bool prefiltered)
{
    if (means3D.ndimension() != 2 || means3D.size(1) != 3) AT_ERROR("means3D must have dimensions (num_points, 3)");
    Marshalled m;
    m.bg = f32c(background); m.means3D = f32c(means3D); m.colors = f32c(colors); m.opacity = f32c(opacity);
    m.scales = f32c(scales); m.rotations = f32c(rotations); m.cov3D = f32c(cov3D_precomp); m.view = f32c(viewmatrix);
    m.proj = f32c(projmatrix); m.sh = f32c(sh); m.campos = f32c(campos);
    gsb_raster_args& a = m.a;
    a.P = (int)means3D.size(0);
    a.D = degree;
    a.M = (m.sh.defined() && m.sh.numel() > 0) ? (int)m.sh.size(1) : 0;
    a.width = image_width; a.height = image_height;
    a.background = fptr(m.bg); a.means3D = fptr(m.means3D); a.shs = fptr(m.sh); a.colors_precomp = fptr(m.colors);
    a.opacities = fptr(m.opacity); a.scales = fptr(m.scales); a.scale_modifier = scale_modifier;
    a.rotations = fptr(m.rotations); a.cov3D_precomp = fptr(m.cov3D); a.viewmatrix = fptr(m.view);
    a.projmatrix = fptr(m.proj); a.cam_pos = fptr(m.campos);
    a.tan_fovx = tan_fovx; a.tan_fovy = tan_fovy; a.prefiltered = prefiltered ? 1 : 0;
    a.tile_row_begin = a.tile_row_end = 0;   // whole image (the tile-row shard is a multi-GPU extension)
    return m;
}

// gsb_alloc_fn over a torch byte tensor (the role of resizeFunctional, src/Rasterizer.cu:127-134).
void* grow(void* user, size_t bytes)
{
    auto* t = static_cast<torch::Tensor*>(user);
    t->resize_({(long long)bytes});
    return t->data_ptr();
}

}  // namespace

std::tuple<int, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
RasterizeGaussiansCUDA(const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& colors,
                       const torch::Tensor& opacity, const torch::Tensor& scales, const torch::Tensor& rotations,
                       const float scale_modifier, const torch::Tensor& cov3D_precomp, const torch::Tensor& viewmatrix,
                       const torch::Tensor& projmatrix, const float tan_fovx, const float tan_fovy, const int image_height,
                       const int image_width, const torch::Tensor& sh, const int degree, const torch::Tensor& campos,
                       const bool prefiltered, const int device_num)
{
    (void)device_num;  // the device is the one the tensors live on
    c10::cuda::CUDAGuard guard(means3D.device());
    Marshalled m = marshal(background, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix,
                           projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree, campos, prefiltered);
    const int P = m.a.P, H = image_height, W = image_width;
    auto fopts = means3D.options().dtype(torch::kFloat32);
    torch::Tensor out_color = torch::empty({3, H, W}, fopts);
    torch::Tensor out_depth = torch::empty({1, H, W}, fopts);
    torch::Tensor radii = torch::empty({P}, means3D.options().dtype(torch::kInt32));
    auto bopts = torch::TensorOptions().dtype(torch::kByte).device(means3D.device());
    torch::Tensor geomBuffer = torch::empty({0}, bopts), binningBuffer = torch::empty({0}, bopts), imgBuffer = torch::empty({0}, bopts);
    int rendered = 0;
    if (P != 0) {
        const int rc = gsb_forward(&m.a, grow, &geomBuffer, grow, &binningBuffer, grow, &imgBuffer, out_color.data_ptr<float>(),
                                   out_depth.data_ptr<float>(), radii.data_ptr<int>(), stream());
        check(rc);
        rendered = rc;
    } else {  // the reference returns its fill values (src/Rasterizer.cu:170-172, :182)
        out_color.zero_();
        out_depth.zero_();
    }
    return std::make_tuple(rendered, out_color, radii, geomBuffer, binningBuffer, imgBuffer, out_depth);
}

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
RasterizeGaussiansBackwardCUDA(const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& radii,
                               const torch::Tensor& colors, const torch::Tensor& scales, const torch::Tensor& rotations,
                               const float scale_modifier, const torch::Tensor& cov3D_precomp, const torch::Tensor& viewmatrix,
                               const torch::Tensor& projmatrix, const float tan_fovx, const float tan_fovy,
                               const torch::Tensor& dL_dout_color, const torch::Tensor& sh, const int degree,
                               const torch::Tensor& campos, const torch::Tensor& geomBuffer, const int R,
                               const torch::Tensor& binningBuffer, const torch::Tensor& imageBuffer)
{
    c10::cuda::CUDAGuard guard(means3D.device());
    const int H = (int)dL_dout_color.size(1), W = (int)dL_dout_color.size(2);
    // opacities are not an input of the backward (rasterizer.h:55-83): they live in the geometry state
    Marshalled m = marshal(background, means3D, colors, torch::Tensor(), scales, rotations, scale_modifier,
                           cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, H, W, sh, degree, campos, false);
    const int P = m.a.P, M = m.a.M;
    auto fopts = means3D.options().dtype(torch::kFloat32);
    torch::Tensor dL_dmeans3D = torch::empty({P, 3}, fopts), dL_dmeans2D = torch::empty({P, 3}, fopts);
    torch::Tensor dL_dcolors = torch::empty({P, 3}, fopts), dL_dconic = torch::empty({P, 2, 2}, fopts);
    torch::Tensor dL_dopacity = torch::empty({P, 1}, fopts), dL_dcov3D = torch::empty({P, 6}, fopts);
    torch::Tensor dL_dsh = torch::empty({P, M, 3}, fopts), dL_dscales = torch::empty({P, 3}, fopts);
    torch::Tensor dL_drotations = torch::empty({P, 4}, fopts);
    if (P != 0) {
        torch::Tensor dpix = dL_dout_color.to(torch::kFloat32).contiguous();
        gsb_grad_outputs g;
        g.dL_dmean2D = dL_dmeans2D.data_ptr<float>(); g.dL_dconic = dL_dconic.data_ptr<float>();
        g.dL_dopacity = dL_dopacity.data_ptr<float>(); g.dL_dcolor = dL_dcolors.data_ptr<float>();
        g.dL_dmean3D = dL_dmeans3D.data_ptr<float>(); g.dL_dcov3D = dL_dcov3D.data_ptr<float>();
        g.dL_dsh = M ? dL_dsh.data_ptr<float>() : nullptr;
        g.dL_dscale = m.a.scales ? dL_dscales.data_ptr<float>() : nullptr;
        g.dL_drot = m.a.rotations ? dL_drotations.data_ptr<float>() : nullptr;
        check(gsb_backward(&m.a, R, radii.data_ptr<int>(), geomBuffer.data_ptr(), binningBuffer.data_ptr(), imageBuffer.data_ptr(),
                           dpix.data_ptr<float>(), &g, stream()));
        if (!m.a.scales) dL_dscales.zero_();
        if (!m.a.rotations) dL_drotations.zero_();
    }
    return std::make_tuple(dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations);
}

torch::Tensor markVisible(torch::Tensor& means3D, torch::Tensor& viewmatrix, torch::Tensor& projmatrix)
{
    c10::cuda::CUDAGuard guard(means3D.device());
    const int P = (int)means3D.size(0);
    torch::Tensor present = torch::empty({P}, means3D.options().dtype(torch::kBool));
    if (P > 0) {
        torch::Tensor m = f32c(means3D), v = f32c(viewmatrix), p = f32c(projmatrix);
        check(gsb_mark_visible(P, fptr(m), fptr(v), fptr(p), (uint8_t*)present.data_ptr<bool>(), stream()));
    }
    return present;
}

torch::Tensor RasterizeGaussiansfilterCUDA(const torch::Tensor& means3D, const torch::Tensor& scales,
                                           const torch::Tensor& rotations, const float scale_modifier,
                                           const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix,
                                           const float tan_fovx, const float tan_fovy, const int image_height,
                                           const int image_width, const bool prefiltered, int device_num)
{
    (void)device_num;
    c10::cuda::CUDAGuard guard(means3D.device());
    torch::Tensor none;
    Marshalled m = marshal(none, means3D, none, none, scales, rotations, scale_modifier, none, viewmatrix, projmatrix, tan_fovx,
                           tan_fovy, image_height, image_width, none, 0, none, prefiltered);
    torch::Tensor radii = torch::empty({m.a.P}, means3D.options().dtype(torch::kInt32));
    if (m.a.P != 0) check(gsb_visible_filter(&m.a, radii.data_ptr<int>(), stream()));
    return radii;
}

torch::Tensor filter_radii(torch::Tensor means3D, torch::Tensor scales, torch::Tensor rotations, int device_num,
                           GaussianRasterizationSettings rs)
{
    torch::NoGradGuard no_grad;
    return RasterizeGaussiansfilterCUDA(means3D, scales, rotations, rs.scale_modifier, rs.viewmatrix, rs.projmatrix, rs.tanfovx,
                                        rs.tanfovy, rs.image_height, rs.image_width, rs.prefiltered, device_num);
}

torch::autograd::tensor_list rasterize_gaussians(torch::Tensor means3D, torch::Tensor means2D, torch::Tensor sh,
                                                 torch::Tensor colors_precomp, torch::Tensor opacities, torch::Tensor scales,
                                                 torch::Tensor rotations, torch::Tensor cov3Ds_precomp, int device_num,
                                                 GaussianRasterizationSettings raster_settings)
{
    return _RasterizeGaussians::apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                                      raster_settings, device_num);
}

torch::autograd::tensor_list _RasterizeGaussians::forward(torch::autograd::AutogradContext* ctx, torch::Tensor means3D,
                                                          torch::Tensor means2D, torch::Tensor sh, torch::Tensor colors_precomp,
                                                          torch::Tensor opacities, torch::Tensor scales, torch::Tensor rotations,
                                                          torch::Tensor cov3Ds_precomp, GaussianRasterizationSettings rs,
                                                          int device_num)
{
    (void)means2D;
    auto [num_rendered, color, radii, geomBuffer, binningBuffer, imgBuffer, depth] = RasterizeGaussiansCUDA(
        rs.bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier, cov3Ds_precomp, rs.viewmatrix,
        rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height, rs.image_width, sh, rs.sh_degree, rs.camera_center,
        rs.prefiltered, device_num);
    ctx->saved_data["num_rendered"] = num_rendered;
    ctx->saved_data["scale_modifier"] = (double)rs.scale_modifier;
    ctx->saved_data["tanfovx"] = (double)rs.tanfovx;
    ctx->saved_data["tanfovy"] = (double)rs.tanfovy;
    ctx->saved_data["sh_degree"] = rs.sh_degree;
    auto keep = [](const torch::Tensor& t) { return t.defined() ? t : torch::Tensor(); };
    ctx->save_for_backward({rs.bg, means3D, radii, keep(colors_precomp), keep(scales), keep(rotations), keep(cov3Ds_precomp),
                            rs.viewmatrix, rs.projmatrix, keep(sh), rs.camera_center, geomBuffer, binningBuffer, imgBuffer});
    ctx->mark_non_differentiable({radii, depth});
    return {color, radii, depth};
}

torch::autograd::tensor_list _RasterizeGaussians::backward(torch::autograd::AutogradContext* ctx,
                                                           torch::autograd::tensor_list grad_outputs)
{
    // only d(color) is consumed (Rasterizer.cuh:210): radii and the median depth carry no gradient
    auto s = ctx->get_saved_variables();
    auto [g_means2D, g_colors, g_opacity, g_means3D, g_cov3D, g_sh, g_scales, g_rotations] = RasterizeGaussiansBackwardCUDA(
        s[0], s[1], s[2], s[3], s[4], s[5], (float)ctx->saved_data["scale_modifier"].toDouble(), s[6], s[7], s[8],
        (float)ctx->saved_data["tanfovx"].toDouble(), (float)ctx->saved_data["tanfovy"].toDouble(), grad_outputs[0], s[9],
        (int)ctx->saved_data["sh_degree"].toInt(), s[10], s[11], (int)ctx->saved_data["num_rendered"].toInt(), s[12], s[13]);
    auto opt = [](const torch::Tensor& saved, const torch::Tensor& g) { return (saved.defined() && saved.numel() > 0) ? g : torch::Tensor(); };
    // gradients in the order of forward's arguments (Rasterizer.cuh:259-266); settings / device_num get none
    return {g_means3D, g_means2D, opt(s[9], g_sh), opt(s[3], g_colors), g_opacity, opt(s[4], g_scales), opt(s[5], g_rotations),
            opt(s[6], g_cov3D), torch::Tensor(), torch::Tensor()};
}

// ---- fused five-channel pass -------------------------------------------------------------------------
torch::autograd::tensor_list rasterize_gaussians_fused(torch::Tensor means3D, torch::Tensor colors_precomp,
                                                       torch::Tensor opacities, torch::Tensor scales, torch::Tensor rotations,
                                                       GaussianRasterizationSettings raster_settings, bool z_attached)
{
    return _RasterizeGaussiansFused::apply(means3D, colors_precomp, opacities, scales, rotations, raster_settings, z_attached);
}

torch::autograd::tensor_list _RasterizeGaussiansFused::forward(torch::autograd::AutogradContext* ctx, torch::Tensor means3D,
                                                               torch::Tensor colors_precomp, torch::Tensor opacities,
                                                               torch::Tensor scales, torch::Tensor rotations,
                                                               GaussianRasterizationSettings rs, bool z_attached)
{
    c10::cuda::CUDAGuard guard(means3D.device());
    torch::Tensor none;
    Marshalled m = marshal(rs.bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier, none, rs.viewmatrix,
                           rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height, rs.image_width, none, 0, rs.camera_center,
                           rs.prefiltered);
    const int P = m.a.P, H = rs.image_height, W = rs.image_width;
    auto fopts = means3D.options().dtype(torch::kFloat32);
    auto bopts = torch::TensorOptions().dtype(torch::kByte).device(means3D.device());
    torch::Tensor color = torch::zeros({3, H, W}, fopts), depth_sil = torch::zeros({2, H, W}, fopts);
    torch::Tensor median = torch::zeros({1, H, W}, fopts), radii = torch::zeros({P}, means3D.options().dtype(torch::kInt32));
    torch::Tensor geom, binning, img;
    if (P != 0) {
        // sync-free entry point over caller workspaces; the instance capacity is grown on overflow (one
        // synchronisation to read the count, like the reference's cudaMemcpy at rasterizer_impl.cu:285)
        static thread_local long long capacity_hint = 0;
        long long cap = std::max<long long>(capacity_hint, 4ll * P + 4096);
        geom = torch::empty({(long long)gsb_geometry_bytes(P)}, bopts);
        img = torch::empty({(long long)gsb_image_bytes(W, H)}, bopts);
        while (true) {
            binning = torch::empty({(long long)gsb_binning_bytes(cap)}, bopts);
            check(gsb_forward_fused_ws(&m.a, geom.data_ptr(), geom.numel(), binning.data_ptr(), binning.numel(), cap, img.data_ptr(),
                                       img.numel(), color.data_ptr<float>(), depth_sil.data_ptr<float>(), median.data_ptr<float>(),
                                       radii.data_ptr<int>(), stream()));
            const long long R = gsb_num_rendered(geom.data_ptr(), stream());
            if (R == GSB_ERR_OVERFLOW) { cap *= 2; continue; }
            check(R);
            capacity_hint = std::max(capacity_hint, R + R / 4);
            break;
        }
    }
    ctx->saved_data["scale_modifier"] = (double)rs.scale_modifier;
    ctx->saved_data["tanfovx"] = (double)rs.tanfovx;
    ctx->saved_data["tanfovy"] = (double)rs.tanfovy;
    ctx->saved_data["z_attached"] = z_attached;
    ctx->save_for_backward({rs.bg, means3D, radii, colors_precomp, scales, rotations, rs.viewmatrix, rs.projmatrix,
                            rs.camera_center, geom.defined() ? geom : torch::empty({0}, bopts),
                            binning.defined() ? binning : torch::empty({0}, bopts), img.defined() ? img : torch::empty({0}, bopts)});
    ctx->mark_non_differentiable({median, radii});
    return {color, depth_sil, median, radii};
}

torch::autograd::tensor_list _RasterizeGaussiansFused::backward(torch::autograd::AutogradContext* ctx,
                                                                torch::autograd::tensor_list grad_outputs)
{
    auto s = ctx->get_saved_variables();
    const torch::Tensor &bg = s[0], &means3D = s[1], &radii = s[2], &colors = s[3], &scales = s[4], &rotations = s[5];
    c10::cuda::CUDAGuard guard(means3D.device());
    const int H = (int)grad_outputs[0].size(1), W = (int)grad_outputs[0].size(2);
    torch::Tensor none;
    Marshalled m = marshal(bg, means3D, colors, none, scales, rotations, (float)ctx->saved_data["scale_modifier"].toDouble(), none,
                           s[6], s[7], (float)ctx->saved_data["tanfovx"].toDouble(), (float)ctx->saved_data["tanfovy"].toDouble(), H, W,
                           none, 0, s[8], false);
    const int P = m.a.P;
    auto fopts = means3D.options().dtype(torch::kFloat32);
    torch::Tensor g_means3D = torch::zeros({P, 3}, fopts), g_colors = torch::zeros({P, 3}, fopts);
    torch::Tensor g_opacity = torch::zeros({P, 1}, fopts), g_scales = torch::zeros({P, 3}, fopts), g_rot = torch::zeros({P, 4}, fopts);
    if (P != 0) {
        auto dense = [&](const torch::Tensor& g, int ch) {   // an output the loss did not touch has an undefined gradient
            return g.defined() ? g.to(torch::kFloat32).contiguous() : torch::zeros({ch, H, W}, fopts);
        };
        torch::Tensor dC = dense(grad_outputs[0], 3), dD = dense(grad_outputs[1], 2);
        torch::Tensor side = torch::empty({P, 13}, fopts), g_z = torch::empty({P}, fopts);
        gsb_grad_outputs g;
        float* sp = side.data_ptr<float>();
        g.dL_dmean2D = sp; g.dL_dconic = sp + 3ll * P; g.dL_dcov3D = sp + 7ll * P; g.dL_dsh = nullptr;
        g.dL_dopacity = g_opacity.data_ptr<float>(); g.dL_dcolor = g_colors.data_ptr<float>();
        g.dL_dmean3D = g_means3D.data_ptr<float>(); g.dL_dscale = g_scales.data_ptr<float>(); g.dL_drot = g_rot.data_ptr<float>();
        check(gsb_backward_fused(&m.a, radii.data_ptr<int>(), s[9].data_ptr(), s[10].data_ptr(), s[11].data_ptr(), dC.data_ptr<float>(),
                                 dD.data_ptr<float>(), &g, g_z.data_ptr<float>(), ctx->saved_data["z_attached"].toBool() ? 1 : 0, stream()));
    }
    // gradients in the order of forward's arguments; settings / flag get none
    return {g_means3D, g_colors, g_opacity, g_scales, g_rot, torch::Tensor(), torch::Tensor()};
}

torch::Tensor distCUDA2(const torch::Tensor& points, torch::Device device)
{
    c10::cuda::CUDAGuard guard(device);
    const int P = (int)points.size(0);
    torch::Tensor pts = points.to(device).to(torch::kFloat32).contiguous();
    torch::Tensor means = torch::zeros({P}, pts.options());
    if (P > 0) {
        const size_t bytes = gsb_knn_workspace_bytes(P);
        torch::Tensor ws = torch::empty({(long long)bytes}, pts.options().dtype(torch::kByte));
        check(gsb_knn_mean_dist2(P, pts.data_ptr<float>(), means.data_ptr<float>(), ws.data_ptr(), bytes, stream()));
    }
    return means;
}

}  // namespace ORB_SLAM2
