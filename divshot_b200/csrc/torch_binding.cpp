// torch_binding.cpp — the libtorch operator surface over the C-ABI.
//
// The reference's closed `gsplatrast` exposes its CUDA rasterizer to the libtorch C++ trainer
// (link line application/diverseshot-cli/premake5.lua:25-26,87-91: c10_cuda, torch_cuda, c10, torch, torch_cpu).
// This is the drop-in for that surface: a torch::CustomClassHolder that owns a dvs_rast context
// (persistent arenas, no per-step allocation) plus a torch::autograd::Function whose backward returns the
// gradients of means3D / scales / rotations / opacity / sh0 / shN.  Nothing here computes: every call forwards
// device pointers and the current CUDA stream to include/dvs_rast.h.  Built with g++ (no nvcc) into
// divshot_b200/lib/libdvs_torch.so; load with torch.classes.load_library or link from C++.
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/custom_class.h>
#include <torch/script.h>

#include <cstring>

#include "dvs_rast.h"

namespace dvs {

using torch::Tensor;

struct Rasterizer : torch::CustomClassHolder {
    dvs_rast_ctx* ctx = nullptr;
    int64_t device = 0;
    std::vector<Tensor> saved;  // the six parameter tensors of the last forward (kept alive for backward)
    int64_t last_h = 0, last_w = 0;
    // The forward state (tile lists, final_T, n_contrib, records, camera) lives in the context, ONE copy: only the most
    // recent forward can be differentiated.  Every forward bumps `generation`; the autograd nodes remember the value they
    // were created with and refuse to run a backward against another forward's state (two views rendered with the same
    // Rasterizer before (loss1 + loss2).backward() would otherwise get silently wrong gradients).  Use one Rasterizer per
    // view that is alive at the same time.
    int64_t generation = 0;
    void check_generation(int64_t g) const {
        TORCH_CHECK(g == generation, "dvs::Rasterizer: backward of forward #", g, " but the context holds the state of forward #",
                    generation, " — a Rasterizer keeps ONE outstanding forward; render concurrent views with separate Rasterizers");
    }

    explicit Rasterizer(int64_t dev) : device(dev) {
        TORCH_CHECK(dvs_rast_create((int)dev, &ctx) == DVS_OK,
                    "dvs_rast_create failed: no usable CUDA device (the rasterizer has no CPU path)");
    }
    ~Rasterizer() override { dvs_rast_destroy(ctx); }

    static const float* fptr(const Tensor& t, const char* name) {
        TORCH_CHECK(t.is_cuda() && t.scalar_type() == torch::kFloat32 && t.is_contiguous(), name,
                    " must be a contiguous CUDA float32 tensor");
        return t.numel() ? t.data_ptr<float>() : nullptr;
    }

    // camera: float32 CPU tensor [48] = view[16] | proj[16] | campos[3] | tanfovx tanfovy | W H | bg[3] |
    //         scale_modifier | sh_degree | sh_rest_alloc | flags | pad
    static dvs_camera unpack_camera(const Tensor& cam) {
        TORCH_CHECK(!cam.is_cuda() && cam.scalar_type() == torch::kFloat32 && cam.numel() >= 47, "camera: CPU float32 [48]");
        const float* c = cam.contiguous().data_ptr<float>();
        dvs_camera d{};
        std::memcpy(d.view, c, 64); std::memcpy(d.proj, c + 16, 64); std::memcpy(d.campos, c + 32, 12);
        d.tanfovx = c[35]; d.tanfovy = c[36]; d.width = (int32_t)c[37]; d.height = (int32_t)c[38];
        std::memcpy(d.bg, c + 39, 12);
        d.scale_modifier = c[42]; d.sh_degree = (int32_t)c[43]; d.sh_rest_alloc = (int32_t)c[44];
        d.flags = (uint32_t)c[45];
        return d;
    }

    std::tuple<Tensor, Tensor> forward(const Tensor& camera, const Tensor& means3D, const Tensor& scales,
                                       const Tensor& quats, const Tensor& opacities, const Tensor& sh0,
                                       const Tensor& shN) {
        c10::cuda::CUDAGuard guard((c10::DeviceIndex)device);
        const dvs_camera cam = unpack_camera(camera);
        const int64_t N = means3D.size(0);
        auto opts = means3D.options();
        Tensor image = torch::empty({3, cam.height, cam.width}, opts);
        Tensor radii = torch::empty({N}, opts.dtype(torch::kInt32));
        dvs_params p{fptr(means3D, "means3D"), fptr(scales, "scales"), fptr(quats, "quats"),
                     fptr(opacities, "opacities"), fptr(sh0, "sh0"), fptr(shN, "shN")};
        auto st = c10::cuda::getCurrentCUDAStream((c10::DeviceIndex)device).stream();
        const int rc = dvs_rast_forward(ctx, &cam, N, &p, image.data_ptr<float>(), radii.data_ptr<int32_t>(), st);
        TORCH_CHECK(rc == DVS_OK, "dvs_rast_forward: ", dvs_rast_last_error(ctx));
        saved = {means3D, scales, quats, opacities, sh0, shN};
        last_h = cam.height; last_w = cam.width;
        generation++;
        return {image, radii};
    }

    std::vector<Tensor> backward(const Tensor& dL_dpix, int64_t flags) {
        c10::cuda::CUDAGuard guard((c10::DeviceIndex)device);
        TORCH_CHECK(saved.size() == 6, "backward without forward");
        std::vector<Tensor> g;
        for (auto& t : saved) g.push_back(torch::empty_like(t));
        Tensor mean2D = torch::empty({saved[0].size(0), 2}, saved[0].options());
        dvs_params p{fptr(saved[0], "means3D"), fptr(saved[1], "scales"), fptr(saved[2], "quats"),
                     fptr(saved[3], "opacities"), fptr(saved[4], "sh0"), fptr(saved[5], "shN")};
        dvs_grads gr{};
        gr.means3D = g[0].data_ptr<float>(); gr.scales = g[1].data_ptr<float>(); gr.quats = g[2].data_ptr<float>();
        gr.opacities = g[3].data_ptr<float>(); gr.sh0 = g[4].data_ptr<float>();
        gr.shN = g[5].numel() ? g[5].data_ptr<float>() : nullptr;
        gr.mean2D = mean2D.data_ptr<float>();
        Tensor dl = dL_dpix.contiguous();
        auto st = c10::cuda::getCurrentCUDAStream((c10::DeviceIndex)device).stream();
        const int rc = dvs_rast_backward(ctx, &p, fptr(dl, "dL_dpix"), &gr, (uint32_t)flags, st);
        TORCH_CHECK(rc == DVS_OK, "dvs_rast_backward: ", dvs_rast_last_error(ctx));
        g.push_back(mean2D);  // screen-space gradient for the trainer's densification statistics
        return g;
    }

    // Row F4 (dvs_rast_forward_aux / dvs_rast_backward_aux): depth / alpha [2,H,W] and normal [3,H,W] maps of the last forward
    std::tuple<Tensor, Tensor> forward_aux() {
        c10::cuda::CUDAGuard guard((c10::DeviceIndex)device);
        TORCH_CHECK(saved.size() == 6, "forward_aux without forward");
        TORCH_CHECK(last_h > 0 && last_w > 0, "forward_aux without forward");
        auto opts = saved[0].options();
        Tensor aux = torch::empty({2, last_h, last_w}, opts), normal = torch::empty({3, last_h, last_w}, opts);
        dvs_params p{fptr(saved[0], "means3D"), fptr(saved[1], "scales"), fptr(saved[2], "quats"),
                     fptr(saved[3], "opacities"), fptr(saved[4], "sh0"), fptr(saved[5], "shN")};
        auto st = c10::cuda::getCurrentCUDAStream((c10::DeviceIndex)device).stream();
        const int rc = dvs_rast_forward_aux(ctx, &p, aux.data_ptr<float>(), normal.data_ptr<float>(), st);
        TORCH_CHECK(rc == DVS_OK, "dvs_rast_forward_aux: ", dvs_rast_last_error(ctx));
        return {aux, normal};
    }
    // gradients of <image, dL_dpix> + <depth/alpha, dL_daux> + <normal, dL_dnormal>; an undefined / empty tensor = that loss is absent
    std::vector<Tensor> backward_aux(const Tensor& dL_dpix, const Tensor& dL_daux, const Tensor& dL_dnormal, int64_t flags) {
        c10::cuda::CUDAGuard guard((c10::DeviceIndex)device);
        TORCH_CHECK(saved.size() == 6, "backward without forward");
        std::vector<Tensor> g;
        for (auto& t : saved) g.push_back(torch::empty_like(t));
        Tensor mean2D = torch::empty({saved[0].size(0), 2}, saved[0].options());
        dvs_params p{fptr(saved[0], "means3D"), fptr(saved[1], "scales"), fptr(saved[2], "quats"),
                     fptr(saved[3], "opacities"), fptr(saved[4], "sh0"), fptr(saved[5], "shN")};
        dvs_grads gr{};
        gr.means3D = g[0].data_ptr<float>(); gr.scales = g[1].data_ptr<float>(); gr.quats = g[2].data_ptr<float>();
        gr.opacities = g[3].data_ptr<float>(); gr.sh0 = g[4].data_ptr<float>();
        gr.shN = g[5].numel() ? g[5].data_ptr<float>() : nullptr;
        gr.mean2D = mean2D.data_ptr<float>();
        Tensor dl = dL_dpix.contiguous();
        Tensor da = dL_daux.defined() && dL_daux.numel() ? dL_daux.contiguous() : Tensor();
        Tensor dn = dL_dnormal.defined() && dL_dnormal.numel() ? dL_dnormal.contiguous() : Tensor();
        auto st = c10::cuda::getCurrentCUDAStream((c10::DeviceIndex)device).stream();
        const int rc = dvs_rast_backward_aux(ctx, &p, fptr(dl, "dL_dpix"), da.defined() ? fptr(da, "dL_daux") : nullptr,
                                             dn.defined() ? fptr(dn, "dL_dnormal") : nullptr, &gr, (uint32_t)flags, st);
        TORCH_CHECK(rc == DVS_OK, "dvs_rast_backward_aux: ", dvs_rast_last_error(ctx));
        g.push_back(mean2D);
        return g;
    }

    std::vector<int64_t> stats() const {
        dvs_stats s{};
        dvs_rast_get_stats(ctx, &s);
        return {s.num_gaussians, s.num_visible, s.num_dups, s.max_tile_len, s.tiles_x, s.tiles_y};
    }
};

// autograd bridge: image, radii = dvs::rasterize(rasterizer, camera, means3D, scales, quats, opacities, sh0, shN)
struct RasterizeFn : torch::autograd::Function<RasterizeFn> {
    static torch::autograd::variable_list forward(torch::autograd::AutogradContext* actx,
                                                  const c10::intrusive_ptr<Rasterizer>& r, const Tensor& camera,
                                                  const Tensor& means3D, const Tensor& scales, const Tensor& quats,
                                                  const Tensor& opacities, const Tensor& sh0, const Tensor& shN) {
        auto out = r->forward(camera, means3D.contiguous(), scales.contiguous(), quats.contiguous(),
                              opacities.contiguous(), sh0.contiguous(), shN.contiguous());
        actx->saved_data["rast"] = r;
        actx->saved_data["generation"] = r->generation;
        actx->mark_non_differentiable({std::get<1>(out)});
        return {std::get<0>(out), std::get<1>(out)};
    }
    static torch::autograd::variable_list backward(torch::autograd::AutogradContext* actx,
                                                   torch::autograd::variable_list grad_out) {
        auto r = actx->saved_data["rast"].toCustomClass<Rasterizer>();
        r->check_generation(actx->saved_data["generation"].toInt());
        auto g = r->backward(grad_out[0], 0);
        return {Tensor(), Tensor(), g[0], g[1], g[2], g[3], g[4], g[5]};
    }
};

// autograd bridge with the auxiliary maps: image, radii, depth_alpha [2,H,W], normal [3,H,W] = dvs::rasterize_aux(...)
struct RasterizeAuxFn : torch::autograd::Function<RasterizeAuxFn> {
    static torch::autograd::variable_list forward(torch::autograd::AutogradContext* actx,
                                                  const c10::intrusive_ptr<Rasterizer>& r, const Tensor& camera,
                                                  const Tensor& means3D, const Tensor& scales, const Tensor& quats,
                                                  const Tensor& opacities, const Tensor& sh0, const Tensor& shN) {
        auto out = r->forward(camera, means3D.contiguous(), scales.contiguous(), quats.contiguous(),
                              opacities.contiguous(), sh0.contiguous(), shN.contiguous());
        auto aux = r->forward_aux();
        actx->saved_data["rast"] = r;
        actx->saved_data["generation"] = r->generation;
        actx->mark_non_differentiable({std::get<1>(out)});
        return {std::get<0>(out), std::get<1>(out), std::get<0>(aux), std::get<1>(aux)};
    }
    static torch::autograd::variable_list backward(torch::autograd::AutogradContext* actx,
                                                   torch::autograd::variable_list grad_out) {
        auto r = actx->saved_data["rast"].toCustomClass<Rasterizer>();
        r->check_generation(actx->saved_data["generation"].toInt());
        Tensor dpix = grad_out[0].defined() ? grad_out[0] : torch::zeros({3, r->last_h, r->last_w}, r->saved[0].options());
        auto g = r->backward_aux(dpix, grad_out[2], grad_out[3], 0);
        return {Tensor(), Tensor(), g[0], g[1], g[2], g[3], g[4], g[5]};
    }
};

std::tuple<Tensor, Tensor, Tensor, Tensor> rasterize_aux(const c10::intrusive_ptr<Rasterizer>& r, const Tensor& camera,
                                                         const Tensor& means3D, const Tensor& scales, const Tensor& quats,
                                                         const Tensor& opacities, const Tensor& sh0, const Tensor& shN) {
    auto out = RasterizeAuxFn::apply(r, camera, means3D, scales, quats, opacities, sh0, shN);
    return {out[0], out[1], out[2], out[3]};
}

std::tuple<Tensor, Tensor> rasterize(const c10::intrusive_ptr<Rasterizer>& r, const Tensor& camera,
                                     const Tensor& means3D, const Tensor& scales, const Tensor& quats,
                                     const Tensor& opacities, const Tensor& sh0, const Tensor& shN) {
    auto out = RasterizeFn::apply(r, camera, means3D, scales, quats, opacities, sh0, shN);
    return {out[0], out[1]};
}

TORCH_LIBRARY(dvs, m) {
    m.class_<Rasterizer>("Rasterizer")
        .def(torch::init<int64_t>())
        .def("forward", &Rasterizer::forward)
        .def("backward", &Rasterizer::backward)
        .def("forward_aux", &Rasterizer::forward_aux)
        .def("backward_aux", &Rasterizer::backward_aux)
        .def("stats", &Rasterizer::stats);
    m.def("rasterize(__torch__.torch.classes.dvs.Rasterizer r, Tensor camera, Tensor means3D, Tensor scales, "
          "Tensor quats, Tensor opacities, Tensor sh0, Tensor shN) -> (Tensor, Tensor)");
    m.def("rasterize_aux(__torch__.torch.classes.dvs.Rasterizer r, Tensor camera, Tensor means3D, Tensor scales, "
          "Tensor quats, Tensor opacities, Tensor sh0, Tensor shN) -> (Tensor, Tensor, Tensor, Tensor)");
}
TORCH_LIBRARY_IMPL(dvs, Autograd, m) {
    m.impl("rasterize", &rasterize);
    m.impl("rasterize_aux", &rasterize_aux);
}
TORCH_LIBRARY_IMPL(dvs, CUDA, m) {
    m.impl("rasterize", &rasterize);
    m.impl("rasterize_aux", &rasterize_aux);
}

}  // namespace dvs
