// host_capi.cpp — C entry points over the C++ host (RenderGraph + HybridRenderPath) for harnesses that are not C++
// (tests, bench.py). A C++ maintainer of the reference uses the classes directly; see INTEGRATION.md.
#include <string.h>

#include <memory>

#include "hybrid_render_path.h"

struct vhrh_renderer {
    std::unique_ptr<ResourceManager> resource_manager;
    std::unique_ptr<RenderGraph> render_graph;
    std::unique_ptr<HybridRenderPath> path;
    std::string error;
    bool built = false;
};

namespace {
thread_local std::string g_host_error;
template <typename F>
int guarded(vhrh_renderer *r, F &&f) {
    try {
        f();
        return VHR_OK;
    } catch (const VhrHostError &e) {
        g_host_error = e.message;
        if (r) r->error = e.message;
        return e.status;
    } catch (const std::exception &e) {
        g_host_error = e.what();
        return VHR_ERR_INVALID;
    }
}
}  // namespace

extern "C" {

const char *vhrh_last_error(void) { return g_host_error.c_str(); }

// Renderer::Renderer (renderer.cpp:18-44) minus window / swapchain / UI
int vhrh_renderer_create(int device, void *cuda_stream, uint32_t width, uint32_t height, vhrh_renderer **out) {
    if (!out) return VHR_ERR_INVALID;
    *out = nullptr;
    auto r = std::make_unique<vhrh_renderer>();
    int rc = guarded(r.get(), [&] {
        r->resource_manager = std::make_unique<ResourceManager>(device, cuda_stream, width, height);
        r->render_graph = std::make_unique<RenderGraph>(*r->resource_manager);
        r->path = std::make_unique<HybridRenderPath>(*r->render_graph, *r->resource_manager);
    });
    if (rc == VHR_OK) *out = r.release();
    return rc;
}

void vhrh_renderer_destroy(vhrh_renderer *r) {
    if (!r) return;
    guarded(r, [&] {
        if (r->built) r->path->DeregisterPath(*r->render_graph, *r->resource_manager);
    });
    delete r;
}

vhr_context *vhrh_context(vhrh_renderer *r) { return r ? r->resource_manager->ctx : nullptr; }

// SceneLoader::LoadScene -> ResourceManager::UpdateGeometry (scene_loader.cpp:331): flat arrays, one mesh per
// `prims_per_mesh` primitives (0 = a single mesh); the flat primitive order is what matters (object ids).
int vhrh_load_scene(vhrh_renderer *r, const void *vertices, uint32_t n_vertices, const uint32_t *indices, uint32_t n_indices,
                    const void *primitives, uint32_t n_primitives, uint32_t prims_per_mesh) {
    if (!r) return VHR_ERR_INVALID;
    return guarded(r, [&] {
        std::vector<Vertex> v((const Vertex *)vertices, (const Vertex *)vertices + n_vertices);
        std::vector<uint32_t> idx(indices, indices + n_indices);
        Scene scene;
        const Primitive *p = (const Primitive *)primitives;
        uint32_t per = prims_per_mesh ? prims_per_mesh : (n_primitives ? n_primitives : 1);
        for (uint32_t i = 0; i < n_primitives; i += per) {
            Mesh m;
            for (uint32_t k = i; k < n_primitives && k < i + per; ++k) m.primitives.push_back(p[k]);
            scene.meshes.push_back(std::move(m));
        }
        r->resource_manager->UpdateGeometry(v, idx, scene);
    });
}

// The ImGui radio buttons of HybridRenderPath::ImGuiDrawSettings (hybrid_render_path.cpp:394-441) + Rebuild().
int vhrh_set_modes(vhrh_renderer *r, int shadow_mode, int ambient_occlusion_mode, int reflection_mode, int denoise, int svgf_fused) {
    if (!r) return VHR_ERR_INVALID;
    return guarded(r, [&] {
        r->path->shadow_mode = shadow_mode;
        r->path->ambient_occlusion_mode = ambient_occlusion_mode;
        r->path->reflection_mode = reflection_mode;
        r->path->denoise_shadow_and_ao = denoise != 0;
        r->path->svgf_fused = svgf_fused != 0;
        if (r->built) r->path->Rebuild(); else r->path->Build();
        r->built = true;
    });
}

// gbuffer_mode 0: the harness has uploaded the G-buffer images itself; 1: run the CUDA primary-ray producer
int vhrh_set_gbuffer_producer(vhrh_renderer *r, int gbuffer_mode) {
    if (!r) return VHR_ERR_INVALID;
    return guarded(r, [&] {
        ResourceManager *rm = r->resource_manager.get();
        if (gbuffer_mode == 1)
            r->render_graph->SetGraphicsPassHook("G-Buffer Pass", [rm](vhr_context *ctx) { VHR_CHECK(vhr_gbuffer_pass(ctx, rm->width, rm->height)); });
        else
            r->render_graph->SetGraphicsPassHook("G-Buffer Pass", [](vhr_context *) {});
    });
}

// Renderer::Render (renderer.cpp:184-235): UBO update + RenderGraph::Execute; gather = GatherPerformanceStatistics
int vhrh_render(vhrh_renderer *r, const void *per_frame_data, size_t size, int gather_statistics) {
    if (!r || !per_frame_data || size != sizeof(PerFrameData)) return VHR_ERR_INVALID;
    return guarded(r, [&] {
        VHR_ASSERT(r->built, "vhrh_set_modes must be called before rendering");
        PerFrameData pfd;
        memcpy(&pfd, per_frame_data, sizeof(pfd));
        r->resource_manager->UpdatePerFrameUBO(0, pfd);
        r->render_graph->Execute(0);
        if (gather_statistics) r->render_graph->GatherPerformanceStatistics();
    });
}

uint32_t vhrh_execution_order(vhrh_renderer *r, char *buffer, size_t capacity) {
    // passes joined by '\n'; returns the number of passes
    if (!r) return 0;
    std::string s;
    for (const std::string &p : r->render_graph->execution_order) { if (!s.empty()) s += '\n'; s += p; }
    if (buffer && capacity) { strncpy(buffer, s.c_str(), capacity - 1); buffer[capacity - 1] = 0; }
    return (uint32_t)r->render_graph->execution_order.size();
}

// last = 1: last frame's time, 0: the 0.95/0.05 moving average (render_graph.cpp:199); < 0 if unknown
double vhrh_pass_time_ms(vhrh_renderer *r, const char *pass_name, int last) {
    if (!r || !pass_name) return -1.0;
    auto &m = last ? r->render_graph->last_pass_ms : r->render_graph->pass_timestamps;
    auto it = m.find(pass_name);
    return it == m.end() ? -1.0 : it->second;
}

int vhrh_svgf_push_constants(vhrh_renderer *r, void *out24) {
    if (!r || !out24) return VHR_ERR_INVALID;
    memcpy(out24, &r->path->svgf_push_constants, sizeof(SVGFPushConstants));
    return VHR_OK;
}

}  // extern "C"
