// adapter/pybind.cc -- TEST HARNESS ONLY: exposes the libtorch adapter (adapter/Rasterizer.{cuh,cc}, the
// files a GSORB-SLAM maintainer drops into include/ and src/) to Python so the parity tests can drive
// the C++ surface Render.cc calls -- GaussianRasterizer::forward / Visable / mark_visible, distCUDA2 and
// the fused fast path -- against the ctypes path and the reference fixtures.
#include <torch/extension.h>

#include "Exchange.h"
#include "Rasterizer.cuh"

using namespace ORB_SLAM2;

static GaussianRasterizationSettings settings(int H, int W, double tanx, double tany, torch::Tensor bg, double scale_modifier,
                                              torch::Tensor view, torch::Tensor proj, int sh_degree, torch::Tensor campos)
{
    GaussianRasterizationSettings rs;
    rs.image_height = H; rs.image_width = W; rs.tanfovx = (float)tanx; rs.tanfovy = (float)tany; rs.bg = bg;
    rs.scale_modifier = (float)scale_modifier; rs.viewmatrix = view; rs.projmatrix = proj; rs.sh_degree = sh_degree;
    rs.camera_center = campos; rs.prefiltered = false;
    return rs;
}
static torch::Tensor opt(const c10::optional<torch::Tensor>& t) { return t.has_value() ? *t : torch::Tensor(); }

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m)
{
    m.def("forward", [](torch::Tensor means3D, torch::Tensor means2D, torch::Tensor opacities, c10::optional<torch::Tensor> shs,
                        c10::optional<torch::Tensor> colors, c10::optional<torch::Tensor> scales, c10::optional<torch::Tensor> rotations,
                        c10::optional<torch::Tensor> cov3D, int H, int W, double tanx, double tany, torch::Tensor bg,
                        double scale_modifier, torch::Tensor view, torch::Tensor proj, int sh_degree, torch::Tensor campos) {
        GaussianRasterizer r(settings(H, W, tanx, tany, bg, scale_modifier, view, proj, sh_degree, campos));
        return r.forward(means3D, means2D, opacities, opt(shs), opt(colors), opt(scales), opt(rotations), opt(cov3D), 0);
    });
    m.def("forward_fused", [](torch::Tensor means3D, torch::Tensor colors, torch::Tensor opacities, torch::Tensor scales,
                              torch::Tensor rotations, int H, int W, double tanx, double tany, torch::Tensor bg, double scale_modifier,
                              torch::Tensor view, torch::Tensor proj, torch::Tensor campos, bool z_attached) {
        return rasterize_gaussians_fused(means3D, colors, opacities, scales, rotations,
                                         settings(H, W, tanx, tany, bg, scale_modifier, view, proj, 0, campos), z_attached);
    });
    m.def("visable", [](torch::Tensor means3D, torch::Tensor opacities, torch::Tensor scales, torch::Tensor rotations, int H, int W,
                        double tanx, double tany, torch::Tensor bg, double scale_modifier, torch::Tensor view, torch::Tensor proj,
                        torch::Tensor campos) {
        GaussianRasterizer r(settings(H, W, tanx, tany, bg, scale_modifier, view, proj, 0, campos));
        return std::get<0>(r.Visable(means3D, opacities, scales, rotations, 0));
    });
    m.def("mark_visible", [](torch::Tensor positions, torch::Tensor view, torch::Tensor proj) {
        GaussianRasterizationSettings rs;
        rs.viewmatrix = view; rs.projmatrix = proj;
        return GaussianRasterizer(rs).mark_visible(positions);
    });
    m.def("dist_cuda2", [](torch::Tensor points) { return distCUDA2(points, points.device()); });
    // the C++ host of the exchange step (adapter/Exchange.h); driven by tools/exchange_probe.py under torchrun
    py::class_<GradientExchange>(m, "GradientExchange")
        .def(py::init([](int64_t capacity_floats, torch::Tensor like, const std::string& group_name) {
            return new GradientExchange(capacity_floats, like.device(), group_name);
        }))
        .def("alloc", &GradientExchange::alloc)
        .def("allreduce", &GradientExchange::allreduce, py::arg("t"), py::arg("use_multicast") = true)
        .def("rank", &GradientExchange::rank)
        .def("world_size", &GradientExchange::world_size)
        .def("has_multicast", &GradientExchange::has_multicast);
}
